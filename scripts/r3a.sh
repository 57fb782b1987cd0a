# round 2, call 3a (1 GPU): register bitonic sort for power-of-two keys per lane (dblp / twitter 7 -> 8, collab 13 -> 16) vs merge path
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_gset.py tests/test_gpu_philox_parity.py tests/test_gpu_statistics.py tests/test_gpu_fullsize.py tests/test_gpu_shard.py -x -q -m gpu --timeout 900 2>&1 | tee gpurun_out/r3a_pytest.log | tail -3
q() { python - "$1" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], "value %.4g ms/step %.3f kernel ms %.3f frac %.4f" % (d["value"], d["ms_per_step"], r["kernel_ms_per_launch"], r["frac"]))
except Exception as e: print(sys.argv[1], "no json", e)
P
}
for rep in 1 2; do
for wl in dblp collab; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --quick > gpurun_out/r3a_${wl}_pow2_$rep.json 2> gpurun_out/r3a_${wl}_pow2_$rep.err; q gpurun_out/r3a_${wl}_pow2_$rep.json
  SUBG_SAMPLER_POW2=0 timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --quick > gpurun_out/r3a_${wl}_odd_$rep.json 2> gpurun_out/r3a_${wl}_odd_$rep.err; q gpurun_out/r3a_${wl}_odd_$rep.json
done
done
timeout 600 python bench.py --workload twitter --steps 3 --warmup 2 --quick > gpurun_out/r3a_twitter_pow2.json 2> gpurun_out/r3a_twitter_pow2.err; q gpurun_out/r3a_twitter_pow2.json
SUBG_SAMPLER_POW2=0 timeout 600 python bench.py --workload twitter --steps 3 --warmup 2 --quick > gpurun_out/r3a_twitter_odd.json 2> gpurun_out/r3a_twitter_odd.err; q gpurun_out/r3a_twitter_odd.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gset_sample -s 1 -c 1 -f -o gpurun_out/r3a_sampler_dblp python bench.py --workload dblp --steps 2 --warmup 1 --quick > gpurun_out/r3a_sampler_dblp.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gset_sample -s 1 -c 1 -f -o gpurun_out/r3a_sampler_collab python bench.py --workload collab --steps 2 --warmup 1 --quick > gpurun_out/r3a_sampler_collab.log 2>&1
