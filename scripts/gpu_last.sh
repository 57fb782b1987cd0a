cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tail -4
python scripts/sampler_bench.py ppa 5 2>&1 | grep -v Warning | tail -1
python scripts/sampler_bench.py dblp 5 2>&1 | grep -v Warning | tail -1
