# round 2, call 3o (1 GPU): 7 walks per lane in flight for 128 < num_walks <= 224 (ppa, collab) vs 8; parity tests
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_philox_parity.py tests/test_gpu_gset.py -x -q -m gpu --timeout 500 2>&1 | tail -2
q() { python - "$1" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], "ms/step %.3f kernel ms %.3f frac %.4f" % (d["ms_per_step"], r["kernel_ms_per_launch"], r["frac"]))
except Exception as e: print(sys.argv[1], "no json", e)
P
}
for rep in 1 2; do for wl in ppa collab; do
  timeout 200 python bench.py --workload $wl --steps 10 --warmup 3 --quick > gpurun_out/r3o_${wl}_gw7_$rep.json 2>/dev/null; q gpurun_out/r3o_${wl}_gw7_$rep.json
  SUBG_SAMPLER_GW8=1 timeout 200 python bench.py --workload $wl --steps 10 --warmup 3 --quick > gpurun_out/r3o_${wl}_gw8_$rep.json 2>/dev/null; q gpurun_out/r3o_${wl}_gw8_$rep.json
done; done
