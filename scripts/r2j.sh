# round 2, call j (1 GPU): lean kernel after the diet: timing on the three LP shapes + ncu source-level capture (ppa)
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_philox_parity.py tests/test_gpu_gset.py -x -q -m gpu --timeout 600 2>&1 | tail -3
run() { w=$1; shift
  env "$@" timeout 300 python bench.py --workload $w --quick --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('$w $*: ms/step %.3f kernel_ms %.3f frac %.3f' % (d['ms_per_step'], r['kernel_ms_per_launch'], r['frac']))"
}
run ppa SUBG_COL_PACK=0
run ppa SUBG_COL_PACK=1
run collab X=1
run dblp X=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gset_sample -s 1 -c 1 -f -o gpurun_out/r2j_lean python bench.py --workload ppa --quick --steps 1 --warmup 1 > gpurun_out/r2j_lean.log 2>&1
ls -la gpurun_out/r2j_lean.ncu-rep
