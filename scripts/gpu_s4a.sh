# session-4 baseline: full GPU suite + default bench + reference arm (short)
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
nproc; nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
( time timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tee gpurun_out/s4a_pytest.log | tail -5 ) 2>&1
SUBG_PROFILE_HOST=1 python bench.py --steps 5 --warmup 3 > gpurun_out/s4a_bench.json 2> gpurun_out/s4a_bench.err
grep -E "gset_sampler|export" gpurun_out/s4a_bench.err | tail -6
cat gpurun_out/s4a_bench.json
