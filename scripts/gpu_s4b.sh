cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_walks.py tests/test_gpu_gset.py -x -q -m gpu --timeout 300 2>&1 | tee gpurun_out/s4b_pytest.log | tail -25
SUBG_PROFILE_HOST=1 python bench.py --steps 5 --warmup 3 --e2e-steps 5 --no-cpu-baseline > gpurun_out/s4b_bench.json 2> gpurun_out/s4b_bench.err
grep -E "gset_sampler|export" gpurun_out/s4b_bench.err | tail -8
python -c "
import json; d=json.load(open('gpurun_out/s4b_bench.json')); print(d['value'], d['e2e'])"
