cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gset.py tests/test_gpu_statistics.py tests/test_gpu_shard.py tests/test_gpu_fullsize.py -x -q -m gpu --timeout 300 2>&1 | tee gpurun_out/s4c_pytest.log | tail -8
python bench.py --steps 5 --warmup 3 --e2e-steps 3 --no-cpu-baseline > gpurun_out/s4c_bench.json 2> gpurun_out/s4c_bench.err
python -c "
import json; d=json.load(open('gpurun_out/s4c_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'], d['e2e']['value'])"
for w in collab dblp; do python scripts/sampler_bench.py $w 3 2>&1 | grep -v Warning | tail -2; done
