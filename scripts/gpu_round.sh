# usage (GPU box via gpurun): bash scripts/gpu_round.sh <tag> -- GPU tests, then a bench line
cd ${GRAFT_REPO_ROOT:-.}
tag=${1:-run}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
timeout 1200 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tee gpurun_out/${tag}_pytest.log | tail -25
timeout 900 python bench.py "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; tail -8 gpurun_out/${tag}_bench.err; cat gpurun_out/${tag}_bench.json
