# round 2, call 3f (1 GPU): gpu suite after the compaction-policy change; Hits@50 acceptance (mean pooling) with the compiled
# reference run at nthread = -1 as a fourth source; bench lines of the shapes the policy touches
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 900 2>&1 | tee gpurun_out/r3f_pytest.log | tail -4
timeout 600 python -u tests/acceptance_hits50.py --aggr mean --steps 400 > gpurun_out/r3f_hits50_mean.txt 2> gpurun_out/r3f_hits50_mean.err
echo "hits50 rc=$?"; tail -12 gpurun_out/r3f_hits50_mean.txt
q() { python - "$1" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], "value %.4g ms/step %.3f kernel ms %.3f frac %.4f build %.3f gather %s" % (d["value"], d["ms_per_step"], r["kernel_ms_per_launch"], r["frac"], r["spg_build_ms_per_step"], (r.get("gather") or {}).get("frac_of_gather_rate")))
except Exception as e: print(sys.argv[1], "no json", e)
P
}
for wl in collab dblp ppa; do timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --quick > gpurun_out/r3f_$wl.json 2> gpurun_out/r3f_$wl.err; q gpurun_out/r3f_$wl.json; done
timeout 600 python bench.py --workload twitter --steps 3 --warmup 2 --quick > gpurun_out/r3f_twitter.json 2> gpurun_out/r3f_twitter.err; q gpurun_out/r3f_twitter.json
