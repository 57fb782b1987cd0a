# round 2, call u (1 GPU): PPR fast kernel v3 (node-only queue, 6 CTAs per SM), two-lane JoinStream
# sampler with the rank bitmap's shared memory released
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu --timeout 900 2>&1 | tee gpurun_out/r2u_pytest.log | tail -6
show() { python - "$1" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], "value %.4g ms/step %.3f | kernel ms %.3f frac %.4f e2e %s" % (d["value"], d["ms_per_step"], r["kernel_ms_per_launch"], r["frac"], (d.get("e2e") or {}).get("value")), r.get("pushes_per_s"), r.get("kernel_launches_per_step"))
    for b in d.get("spjoin_batches") or []:
        st=b.get("stream") or {}
        print("   spjoin B", b.get("batch"), b.get("pattern"), "gather %.4g q/s | stream %s q/s ms %s kshare %s" % (b.get("value"), st.get("value"), st.get("ms_per_batch"), st.get("kernel_share_of_batch")))
except Exception as e: print(sys.argv[1], "no json", e)
P
}
timeout 900 python bench.py --workload citation2-ppr --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_ppr.json 2> gpurun_out/r2u_ppr.err; echo "ppr rc=$?"; tail -2 gpurun_out/r2u_ppr.err | cut -c1-300; show gpurun_out/r2u_ppr.json
timeout 600 python bench.py --workload dblp --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_dblp.json 2> gpurun_out/r2u_dblp.err; echo "dblp rc=$?"; tail -2 gpurun_out/r2u_dblp.err | cut -c1-300; show gpurun_out/r2u_dblp.json
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_ppa.json 2> gpurun_out/r2u_ppa.err; echo "ppa rc=$?"; tail -2 gpurun_out/r2u_ppa.err | cut -c1-300; show gpurun_out/r2u_ppa.json
BA="--workload citation2-ppr --scale 0.25 --steps 1 --warmup 1 --quick"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ppr_push_fast -s 1 -c 1 -f -o gpurun_out/r2u_ppr_fast python bench.py $BA > gpurun_out/r2u_ppr_fast_ncu.log 2>&1
ls -la gpurun_out | grep r2u
SUBG_JOIN_LANES=1 timeout 600 python bench.py --workload dblp --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_dblp_lane1.json 2> gpurun_out/r2u_dblp_lane1.err; echo "dblp one lane rc=$?"; show gpurun_out/r2u_dblp_lane1.json
for P in 256 384; do SUBG_PPR_FAST_P=$P timeout 600 python bench.py --workload citation2-ppr --steps 3 --warmup 3 --quick > gpurun_out/r2u_ppr_P$P.json 2> gpurun_out/r2u_ppr_P$P.err; echo "P=$P"; show gpurun_out/r2u_ppr_P$P.json; done
