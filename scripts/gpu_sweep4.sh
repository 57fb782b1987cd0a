cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
{
for b in 2 3 4 5 6 7; do SUBG_SAMPLER_BLOCKS=$b SUBG_SAMPLER_STOP=1 python scripts/sampler_bench.py ppa 3; done
for b in 2 3 4 5 6 7; do SUBG_SAMPLER_BLOCKS=$b python scripts/sampler_bench.py ppa 3; done
for b in 3 5 8; do SUBG_SAMPLER_BLOCKS=$b python scripts/sampler_bench.py dblp 3; done
for b in 3 5 8; do SUBG_SAMPLER_BLOCKS=$b python scripts/sampler_bench.py collab 3; done
} 2>&1 | grep -v Warning | tee gpurun_out/sweep4.txt
