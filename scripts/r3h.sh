# round 2, call 3h (1 GPU): final state of the round -- gpu test suite, smoke, bench line of every BASELINE workload, reference arm, launch list, ncu
# CPU baseline), the reference arm, launch list + full ncu captures of the dominant kernels
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 900 2>&1 | tee gpurun_out/r3h_pytest.log | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
show() { python - "$1" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d.get("roofline") or {}
    print(sys.argv[1], "value %.4g ms/step %.3f | kernel ms %.3f frac %.4f e2e %s cpu %s" % (d["value"], d["ms_per_step"], r.get("kernel_ms_per_launch") or 0, r.get("frac") or 0, (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value")))
    for b in d.get("spjoin_batches") or []:
        st=b.get("stream") or {}
        print("   spjoin B", b.get("batch"), b.get("pattern"), "gather %.4g q/s | stream %s q/s ms %s kshare %s" % (b.get("value"), st.get("value"), st.get("ms_per_batch"), st.get("kernel_share_of_batch")))
except Exception as e: print(sys.argv[1], "no json", e)
P
}
timeout 900 python bench.py > gpurun_out/r3h_bench_ppa.json 2> gpurun_out/r3h_bench_ppa.err; echo "ppa rc=$?"; tail -2 gpurun_out/r3h_bench_ppa.err | cut -c1-200; show gpurun_out/r3h_bench_ppa.json
timeout 600 python bench.py --impl reference > gpurun_out/r3h_bench_reference.json 2> gpurun_out/r3h_bench_reference.err; echo "reference rc=$?"; cut -c1-400 gpurun_out/r3h_bench_reference.json
for wl in collab dblp citation2-ppr; do
  timeout 900 python bench.py --workload $wl > gpurun_out/r3h_bench_$wl.json 2> gpurun_out/r3h_bench_$wl.err; echo "$wl rc=$?"; tail -2 gpurun_out/r3h_bench_$wl.err | cut -c1-200; show gpurun_out/r3h_bench_$wl.json
done
timeout 900 python bench.py --workload twitter --steps 3 --warmup 3 > gpurun_out/r3h_bench_twitter.json 2> gpurun_out/r3h_bench_twitter.err; echo "twitter rc=$?"; tail -2 gpurun_out/r3h_bench_twitter.err | cut -c1-200; show gpurun_out/r3h_bench_twitter.json
BA="--steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r3h_launches.csv python bench.py $BA > gpurun_out/r3h_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gset_sample -s 1 -c 1 -f -o gpurun_out/r3h_sampler python bench.py $BA --quick > gpurun_out/r3h_sampler.log 2>&1
ls -la gpurun_out | grep r3h
