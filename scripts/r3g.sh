# round 2, call 3g (1 GPU): batch_sampler parity on the device; per-pass times after caching the device's memory size
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_batch_sampler.py tests/test_gpu_gset.py tests/test_gpu_walks.py -x -q -m gpu --timeout 500 2>&1 | tee gpurun_out/r3g_pytest.log | tail -4
q() { python - "$1" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], "value %.4g ms/step %.3f kernel ms %.3f frac %.4f build %.3f" % (d["value"], d["ms_per_step"], r["kernel_ms_per_launch"], r["frac"], r["spg_build_ms_per_step"]))
except Exception as e: print(sys.argv[1], "no json", e)
P
}
for wl in collab dblp ppa; do timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --quick > gpurun_out/r3g_$wl.json 2> gpurun_out/r3g_$wl.err; q gpurun_out/r3g_$wl.json; done
SUBG_COMPACT_SLACK_PCT=0 timeout 300 python bench.py --workload collab --steps 10 --warmup 3 --quick > gpurun_out/r3g_collab_compact.json 2> gpurun_out/r3g_collab_compact.err; q gpurun_out/r3g_collab_compact.json
timeout 600 python bench.py --workload twitter --steps 3 --warmup 2 --quick > gpurun_out/r3g_twitter.json 2> gpurun_out/r3g_twitter.err; q gpurun_out/r3g_twitter.json
