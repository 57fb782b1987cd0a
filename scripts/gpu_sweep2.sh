cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tail -5
M="dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum"
{
python scripts/sampler_bench.py ppa 5
SUBG_L2_FETCH=32 python scripts/sampler_bench.py ppa 5
SUBG_L2_PERSIST=1 python scripts/sampler_bench.py ppa 5
SUBG_L2_FETCH=32 SUBG_L2_PERSIST=1 python scripts/sampler_bench.py ppa 5
python scripts/sampler_bench.py collab 5
python scripts/sampler_bench.py dblp 5
for v in "X=1" "SUBG_L2_FETCH=32" "SUBG_L2_PERSIST=1" "SUBG_SAMPLER_HINTS=0"; do
echo "== ncu $v"
env $v ncu --metrics $M --clock-control none -k regex:gset_sample -s 2 -c 1 python scripts/sampler_bench.py ppa 1 2>&1 | grep -E "dram__|lts__|gpu__time|smsp__inst"
done
} 2>&1 | grep -v Warning | tee gpurun_out/sweep2.txt
