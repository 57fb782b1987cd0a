cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tail -5
{
for c in 0 1; do SUBG_LP_CACHE=$c python scripts/sampler_bench.py ppa 4; done
for b in 5 6 7; do SUBG_SAMPLER_BLOCKS=$b python scripts/sampler_bench.py ppa 4; done
RANKS=1 python scripts/sampler_bench.py ppa 4
for c in 0 1; do SUBG_LP_CACHE=$c python scripts/sampler_bench.py dblp 4; done
for c in 0 1; do SUBG_LP_CACHE=$c python scripts/sampler_bench.py collab 4; done
} 2>&1 | grep -v Warning | tee gpurun_out/sweep7.txt
ncu --set full --clock-control none --import-source on -k regex:gset_sample -s 2 -c 1 -f -o gpurun_out/r1e_sampler python scripts/sampler_bench.py ppa 1 > gpurun_out/r1e_sampler.log 2>&1
