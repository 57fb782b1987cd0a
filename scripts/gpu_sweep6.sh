cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
{
SUBG_PROFILE_HOST=1 python scripts/sampler_bench.py ppa 4
SUBG_PROFILE_HOST=1 python scripts/sampler_bench.py collab 4
} 2>&1 | grep -v Warning | tee gpurun_out/sweep6.txt
