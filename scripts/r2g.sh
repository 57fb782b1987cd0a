# round 2, call g (1 GPU): joiner + hash fix + packed columns
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_spjoin.py tests/test_gpu_philox_parity.py tests/test_gpu_gset.py tests/test_gpu_fullsize.py -x -q -m gpu --timeout 600 2>&1 | tee gpurun_out/r2g_pytest.log | tail -15
run() { # name, env..., workload
  w=$1; shift
  env "$@" timeout 300 python bench.py --workload $w --quick --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('$w $*: ms/step %.3f kernel_ms %.3f frac %.3f' % (d['ms_per_step'], r['kernel_ms_per_launch'], r['frac']))"
}
run ppa SUBG_SAMPLER_HASH=0 SUBG_COL_PACK=0
run ppa SUBG_SAMPLER_HASH=0 SUBG_COL_PACK=1
run ppa SUBG_SAMPLER_HASH=0 SUBG_COL_PACK=1 SUBG_SAMPLER_STOP=1
run ppa SUBG_SAMPLER_HASH=1 SUBG_COL_PACK=1
run ppa SUBG_SAMPLER_HASH=1 SUBG_COL_PACK=1 SUBG_SAMPLER_STOP=6
run ppa SUBG_SAMPLER_HASH=1 SUBG_COL_PACK=1 SUBG_SAMPLER_STOP=1
run ppa SUBG_SAMPLER_HASH=1 SUBG_COL_PACK=1 SUBG_SAMPLER_STOP=2
run collab SUBG_SAMPLER_HASH=0
run collab SUBG_SAMPLER_HASH=1
run collab SUBG_SAMPLER_HASH=1 SUBG_SAMPLER_STOP=2
run dblp SUBG_SAMPLER_HASH=0
run dblp SUBG_SAMPLER_HASH=1
timeout 600 python bench.py --workload dblp --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2g_dblp.json 2> gpurun_out/r2g_dblp.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2g_dblp.json").read().strip().splitlines()[-1])
for b in d["spjoin_batches"]:
    print("dblp spjoin", b.get("batch"), b.get("value"), b.get("ms_per_batch"), b.get("stream"))
P
