# round 2, call 3j (1 GPU): PPR tests after the unweighted-degree shortcut; ppa sampler at 64 registers / 8 resident CTAs
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ppr.py -x -q -m gpu --timeout 500 2>&1 | tail -2
XB=$PWD/surel_plus_b200/_lib/libsubg_b200_xb.so
q() { python - "$1" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], "value %.4g ms/step %.3f kernel ms %.3f frac %.4f" % (d["value"], d["ms_per_step"], r["kernel_ms_per_launch"], r["frac"]))
except Exception as e: print(sys.argv[1], "no json", e)
P
}
for rep in 1 2; do
  timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r3j_ppa_base$rep.json 2> gpurun_out/r3j_ppa_base$rep.err; q gpurun_out/r3j_ppa_base$rep.json
  SUBG_LIB=$XB timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r3j_ppa_r64_7cta$rep.json 2> gpurun_out/r3j_ppa_r64_7cta$rep.err; q gpurun_out/r3j_ppa_r64_7cta$rep.json
  SUBG_LIB=$XB SUBG_SAMPLER_BLOCKS=8 timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r3j_ppa_r64_8cta$rep.json 2> gpurun_out/r3j_ppa_r64_8cta$rep.err; q gpurun_out/r3j_ppa_r64_8cta$rep.json
done
timeout 600 python bench.py --workload citation2-ppr --steps 3 --warmup 3 --no-cpu-baseline --spjoin-batch 1001 > gpurun_out/r3j_ppr.json 2> gpurun_out/r3j_ppr.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r3j_ppr.json").read().strip().splitlines()[-1]); print("ppr value %.4g ms %.1f e2e %s"%(d["value"], d["ms_per_step"], d["e2e"]))
P
