cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spjoin.py tests/test_gpu_ppr.py tests/test_gpu_fullsize.py -x -q -m gpu --timeout 300 2>&1 | tee gpurun_out/s4e_pytest.log | tail -8
python bench.py --steps 5 --warmup 3 --e2e-steps 3 --no-cpu-baseline > gpurun_out/s4e_bench.json 2> gpurun_out/s4e_bench.err
python -c "
import json; d=json.load(open('gpurun_out/s4e_bench.json')); print(d['value'], d['e2e']['value']); print(json.dumps(d['spjoin']))"
