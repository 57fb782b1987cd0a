# round 2, call h (1 GPU): bitonic sort + LP cache in the main sampler kernel; JoinStream host cost
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_spjoin.py tests/test_gpu_philox_parity.py tests/test_gpu_gset.py tests/test_gpu_fullsize.py tests/test_gpu_statistics.py tests/test_gpu_walks.py -x -q -m gpu --timeout 600 2>&1 | tee gpurun_out/r2h_pytest.log | tail -8
run() { w=$1; shift
  env "$@" timeout 300 python bench.py --workload $w --quick --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('$w $*: ms/step %.3f kernel_ms %.3f frac %.3f' % (d['ms_per_step'], r['kernel_ms_per_launch'], r['frac']))"
}
for w in ppa collab dblp; do
run $w SUBG_SAMPLER_HASH=0 SUBG_SORT_MERGE=1 SUBG_LP_CACHE=0 SUBG_COL_PACK=0
run $w SUBG_SAMPLER_HASH=0 SUBG_SORT_MERGE=1 SUBG_LP_CACHE=256 SUBG_COL_PACK=0
run $w SUBG_SAMPLER_HASH=0 SUBG_SORT_MERGE=0 SUBG_LP_CACHE=0 SUBG_COL_PACK=0
run $w SUBG_SAMPLER_HASH=0 SUBG_SORT_MERGE=0 SUBG_LP_CACHE=256 SUBG_COL_PACK=0
run $w SUBG_SAMPLER_HASH=0 SUBG_SORT_MERGE=0 SUBG_LP_CACHE=1024 SUBG_COL_PACK=0
done
run ppa SUBG_SAMPLER_HASH=0 SUBG_SORT_MERGE=0 SUBG_LP_CACHE=256 SUBG_COL_PACK=1
run ppa SUBG_SAMPLER_HASH=0 SUBG_SORT_MERGE=0 SUBG_LP_CACHE=256 SUBG_COL_PACK=1 SUBG_SAMPLER_BLOCKS=6
run ppa SUBG_SAMPLER_HASH=0 SUBG_SORT_MERGE=0 SUBG_LP_CACHE=256 SUBG_COL_PACK=1 SUBG_SAMPLER_BLOCKS=8
timeout 600 python bench.py --workload dblp --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2h_dblp.json 2> gpurun_out/r2h_dblp.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2h_dblp.json").read().strip().splitlines()[-1])
for b in d["spjoin_batches"]:
    st=b.get("stream") or {}
    print("dblp spjoin B", b.get("batch"), "gather q/s %.3g ms %.4f | stream q/s %s ms %s host_us %s kshare %s" % (b.get("value"), b.get("ms_per_batch"), st.get("value"), st.get("ms_per_batch"), st.get("host_us_per_submit"), st.get("kernel_share_of_batch")))
P
