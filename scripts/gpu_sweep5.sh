cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tail -5
{
for b in 4 5 6 7 8; do SUBG_SAMPLER_BLOCKS=$b python scripts/sampler_bench.py ppa 3; done
python scripts/sampler_bench.py ppa 3
SUBG_SAMPLER_STOP=1 python scripts/sampler_bench.py ppa 3
python scripts/sampler_bench.py dblp 3
python scripts/sampler_bench.py collab 3
} 2>&1 | grep -v Warning | tee gpurun_out/sweep5.txt
