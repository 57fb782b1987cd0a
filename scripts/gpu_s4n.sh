cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_gset.py tests/test_gpu_fullsize.py -x -q -m gpu --timeout 600 2>&1 | tee gpurun_out/s4n_pytest.log | tail -8
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s4n_bench.json 2> gpurun_out/s4n_bench.err; echo rc=$?
python -c "
import json; d=json.load(open('gpurun_out/s4n_bench.json')); print(d['value'], d['e2e']['value'], d['spjoin']['value'], d['spjoin']['ms_per_batch'], d['spjoin']['e2e']['value'])"
