# round 2, call q (1 GPU): PPR fast push kernel -- parity tests, citation2 bench fast vs general, ncu of both
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ppr.py -x -q -m gpu --timeout 300 2>&1 | tee gpurun_out/r2q_pytest_ppr.log | tail -8
timeout 1200 python -m pytest tests -x -q -m gpu --timeout 900 2>&1 | tee gpurun_out/r2q_pytest.log | tail -5
show() { python - "$1" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], "value %.4g seeds/s ms/step %.2f | kernel ms %.2f launches/step %s pushes/s %.4g frac %.4f e2e %s" % (d["value"], d["ms_per_step"], r["kernel_ms_per_launch"], r.get("kernel_launches_per_step"), r["pushes_per_s"], r["frac"], (d.get("e2e") or {}).get("value")))
except Exception as e: print("no json", e)
P
}
for f in 0 1; do
  SUBG_PPR_FAST=$f timeout 900 python bench.py --workload citation2-ppr --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_ppr_fast$f.json 2> gpurun_out/r2q_ppr_fast$f.err
  echo "fast=$f rc=$?"; tail -2 gpurun_out/r2q_ppr_fast$f.err | cut -c1-300; show gpurun_out/r2q_ppr_fast$f.json
done
for b in 1 2; do
  SUBG_PPR_BLOCKS=$b timeout 900 python bench.py --workload citation2-ppr --steps 3 --warmup 3 --quick > gpurun_out/r2q_ppr_blocks$b.json 2> gpurun_out/r2q_ppr_blocks$b.err
  echo "blocks=$b rc=$?"; show gpurun_out/r2q_ppr_blocks$b.json
done
BA="--workload citation2-ppr --scale 0.25 --steps 1 --warmup 1 --quick"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2q_ppr_launches.csv python bench.py $BA > gpurun_out/r2q_ppr_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ppr_push_fast -s 1 -c 1 -f -o gpurun_out/r2q_ppr_fast python bench.py $BA > gpurun_out/r2q_ppr_fast_ncu.log 2>&1
SUBG_PPR_FAST=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:ppr_push_kernel -s 1 -c 1 -f -o gpurun_out/r2q_ppr_general python bench.py $BA > gpurun_out/r2q_ppr_general_ncu.log 2>&1
ls -la gpurun_out/ | grep r2q
