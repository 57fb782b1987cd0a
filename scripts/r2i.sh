# round 2, call i (1 GPU): LEAN main kernel; hash kernel with vote-terminated inserts
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_philox_parity.py tests/test_gpu_gset.py tests/test_gpu_fullsize.py tests/test_gpu_shard.py -x -q -m gpu --timeout 600 2>&1 | tee gpurun_out/r2i_pytest.log | tail -8
run() { w=$1; shift
  env "$@" timeout 300 python bench.py --workload $w --quick --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('$w $*: ms/step %.3f kernel_ms %.3f frac %.3f' % (d['ms_per_step'], r['kernel_ms_per_launch'], r['frac']))"
}
for w in ppa collab dblp; do
run $w SUBG_SAMPLER_HASH=0 SUBG_COL_PACK=0
run $w SUBG_SAMPLER_HASH=1 SUBG_COL_PACK=0
run $w SUBG_SAMPLER_HASH=1 SUBG_COL_PACK=0 SUBG_SAMPLER_STOP=6
run $w SUBG_SAMPLER_HASH=1 SUBG_COL_PACK=0 SUBG_SAMPLER_STOP=1
run $w SUBG_SAMPLER_HASH=1 SUBG_COL_PACK=0 SUBG_SAMPLER_STOP=2
done
run ppa SUBG_SAMPLER_HASH=0 SUBG_COL_PACK=1
run ppa SUBG_SAMPLER_HASH=1 SUBG_COL_PACK=0 SUBG_HASH_CAP=1024
run ppa SUBG_SAMPLER_HASH=1 SUBG_COL_PACK=0 SUBG_SAMPLER_BLOCKS=6
run dblp SUBG_SAMPLER_HASH=1 SUBG_SAMPLER_BLOCKS=8
run dblp SUBG_SAMPLER_HASH=1 SUBG_SAMPLER_BLOCKS=10
run collab SUBG_SAMPLER_HASH=1 SUBG_SAMPLER_BLOCKS=8
