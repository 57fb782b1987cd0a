# round 2, call d (1 GPU): hash-dedup sampler kernel: parity tests, then old vs new kernel on the three LP shapes
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_philox_parity.py tests/test_gpu_gset.py tests/test_gpu_fullsize.py tests/test_gpu_shard.py tests/test_gpu_statistics.py -x -q -m gpu --timeout 600 2>&1 | tee gpurun_out/r2d_pytest.log | tail -15
for w in ppa collab dblp; do
  for h in 0 1; do
    SUBG_SAMPLER_HASH=$h timeout 300 python bench.py --workload $w --quick --steps 5 --warmup 3 > gpurun_out/r2d_${w}_h$h.json 2> gpurun_out/r2d_${w}_h$h.err
    python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r2d_${w}_h$h.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("$w hash=$h value %.4g ms/step %.3f kernel_ms %.3f frac %.3f build %.3f" % (d["value"], d["ms_per_step"], r["kernel_ms_per_launch"], r["frac"], r["spg_build_ms_per_step"]))
except Exception as e: print("$w $h failed", e)
P
  done
done
for b in 3 4 5 6; do
  SUBG_SAMPLER_BLOCKS=$b timeout 300 python bench.py --workload ppa --quick --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('ppa blocks=$b kernel_ms %.3f' % r['kernel_ms_per_launch'])"
done
