cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tail -3
{
python scripts/sampler_bench.py ppa 5
SUBG_SAMPLER_BLOCKS=6 python scripts/sampler_bench.py ppa 5
RANKS=1 python scripts/sampler_bench.py ppa 5
python scripts/sampler_bench.py dblp 5
python scripts/sampler_bench.py collab 5
M="smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"
ncu --metrics $M --clock-control none -k regex:gset_sample -s 2 -c 1 python scripts/sampler_bench.py ppa 1 2>&1 | grep -E "smsp__|gpu__time"
} 2>&1 | grep -v Warning | tee gpurun_out/sweep8.txt
