cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spjoin.py tests/test_gpu_gset.py -x -q -m gpu --timeout 300 2>&1 | tail -3
SUBG_PROFILE_HOST=1 python bench.py --steps 5 --warmup 3 > gpurun_out/r1f_bench.json 2> gpurun_out/r1f_bench.err
grep -E "gset_sampler|export:" gpurun_out/r1f_bench.err | tail -8
cat gpurun_out/r1f_bench.json
