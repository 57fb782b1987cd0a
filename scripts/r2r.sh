# round 2, call r (1 GPU): sampler sort with bidirectional merge (default build) vs forward-only merge (libsubg_b200_fwd.so),
# per-batch join launch list at the reference's batch sizes (dblp), ncu of the remaining kernels
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gset.py tests/test_gpu_philox_parity.py tests/test_gpu_fullsize.py -x -q -m gpu --timeout 600 2>&1 | tee gpurun_out/r2r_pytest.log | tail -4
FWD=$PWD/surel_plus_b200/_lib/libsubg_b200_fwd.so
q() { python - "$1" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], "ms/step %.3f kernel ms %.3f frac %.4f" % (d["ms_per_step"], r["kernel_ms_per_launch"], r["frac"]))
except Exception as e: print(sys.argv[1], "no json", e)
P
}
for rep in 1 2; do
for wl in ppa collab dblp; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --quick > gpurun_out/r2r_${wl}_bidir$rep.json 2> gpurun_out/r2r_${wl}_bidir$rep.err; q gpurun_out/r2r_${wl}_bidir$rep.json
  SUBG_LIB=$FWD timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --quick > gpurun_out/r2r_${wl}_fwd$rep.json 2> gpurun_out/r2r_${wl}_fwd$rep.err; q gpurun_out/r2r_${wl}_fwd$rep.json
done
done
# launch list of the dblp bench (JoinStream at B = 2048 triplets: plan vs join kernel)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2r_dblp_launches.csv python bench.py --workload dblp --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r2r_dblp_launches.log 2>&1
# the sampler after the merge change
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gset_sample -s 1 -c 1 -f -o gpurun_out/r2r_sampler python bench.py --steps 2 --warmup 1 --quick > gpurun_out/r2r_sampler.log 2>&1
# SUREL-v1 walks: walk_sample_kernel, rpe_kernel; remap_ids_kernel from the ppa pass
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"walk_sample_kernel|rpe_kernel|walk_join_kernel" -c 3 -f -o gpurun_out/r2r_walks python scripts/walks_bench.py > gpurun_out/r2r_walks.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"remap_ids_kernel|compact_rows_kernel" -s 2 -c 2 -f -o gpurun_out/r2r_remap python bench.py --steps 2 --warmup 1 --quick > gpurun_out/r2r_remap.log 2>&1
ls -la gpurun_out | grep r2r
