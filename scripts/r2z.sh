# round 2, call z (1 GPU): sampler with 4 walks per lane when num_walks <= 128 (dblp, twitter) vs 8; PPR after the value arrays
# moved to the block cache
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gset.py tests/test_gpu_philox_parity.py tests/test_gpu_statistics.py tests/test_gpu_ppr.py -x -q -m gpu --timeout 600 2>&1 | tee gpurun_out/r2z_pytest.log | tail -3
q() { python - "$1" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], "value %.4g ms/step %.3f kernel ms %.3f frac %.4f" % (d["value"], d["ms_per_step"], r["kernel_ms_per_launch"], r["frac"]), r.get("pushes_per_s"))
except Exception as e: print(sys.argv[1], "no json", e)
P
}
for rep in 1 2; do
  timeout 300 python bench.py --workload dblp --steps 10 --warmup 3 --quick > gpurun_out/r2z_dblp_gw4_$rep.json 2> gpurun_out/r2z_dblp_gw4_$rep.err; q gpurun_out/r2z_dblp_gw4_$rep.json
  SUBG_SAMPLER_GW8=1 timeout 300 python bench.py --workload dblp --steps 10 --warmup 3 --quick > gpurun_out/r2z_dblp_gw8_$rep.json 2> gpurun_out/r2z_dblp_gw8_$rep.err; q gpurun_out/r2z_dblp_gw8_$rep.json
done
timeout 600 python bench.py --workload twitter --steps 3 --warmup 2 --quick > gpurun_out/r2z_twitter_gw4.json 2> gpurun_out/r2z_twitter_gw4.err; q gpurun_out/r2z_twitter_gw4.json
SUBG_SAMPLER_GW8=1 timeout 600 python bench.py --workload twitter --steps 3 --warmup 2 --quick > gpurun_out/r2z_twitter_gw8.json 2> gpurun_out/r2z_twitter_gw8.err; q gpurun_out/r2z_twitter_gw8.json
for rep in 1 2; do
  timeout 900 python bench.py --workload citation2-ppr --steps 3 --warmup 3 --quick > gpurun_out/r2z_ppr_$rep.json 2> gpurun_out/r2z_ppr_$rep.err; q gpurun_out/r2z_ppr_$rep.json
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gset_sample -s 1 -c 1 -f -o gpurun_out/r2z_sampler_dblp python bench.py --workload dblp --steps 2 --warmup 1 --quick > gpurun_out/r2z_sampler_dblp.log 2>&1
