# round 2, call v (1 GPU): PPR fast kernel -- eager vs lazy degree fetch, table size; ncu of the sampler on the collab / dblp shapes
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
LAZY=$PWD/surel_plus_b200/_lib/libsubg_b200_lazy.so
show() { python - "$1" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], "value %.4g ms/step %.3f | kernel ms %.3f x %s pushes/s %.4g" % (d["value"], d["ms_per_step"], r["kernel_ms_per_launch"], r.get("kernel_launches_per_step"), r.get("pushes_per_s")))
except Exception as e: print(sys.argv[1], "no json", e)
P
}
SUBG_LIB=$LAZY timeout 300 python -m pytest tests/test_gpu_ppr.py -x -q -m gpu 2>&1 | tail -2
for rep in 1 2; do
  timeout 600 python bench.py --workload citation2-ppr --steps 3 --warmup 3 --quick > gpurun_out/r2v_ppr_eager$rep.json 2> gpurun_out/r2v_ppr_eager$rep.err; show gpurun_out/r2v_ppr_eager$rep.json
  SUBG_LIB=$LAZY timeout 600 python bench.py --workload citation2-ppr --steps 3 --warmup 3 --quick > gpurun_out/r2v_ppr_lazy$rep.json 2> gpurun_out/r2v_ppr_lazy$rep.err; show gpurun_out/r2v_ppr_lazy$rep.json
done
SUBG_PPR_RECORDS=4096 timeout 600 python bench.py --workload citation2-ppr --steps 3 --warmup 3 --quick > gpurun_out/r2v_ppr_eager_R4096.json 2> gpurun_out/r2v_ppr_eager_R4096.err; show gpurun_out/r2v_ppr_eager_R4096.json
SUBG_LIB=$LAZY SUBG_PPR_RECORDS=4096 timeout 600 python bench.py --workload citation2-ppr --steps 3 --warmup 3 --quick > gpurun_out/r2v_ppr_lazy_R4096.json 2> gpurun_out/r2v_ppr_lazy_R4096.err; show gpurun_out/r2v_ppr_lazy_R4096.json
for wl in collab dblp; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gset_sample -s 1 -c 1 -f -o gpurun_out/r2v_sampler_$wl python bench.py --workload $wl --steps 2 --warmup 1 --quick > gpurun_out/r2v_sampler_$wl.log 2>&1
done
ls -la gpurun_out | grep r2v
