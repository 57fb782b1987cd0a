# usage: bash scripts/gpu_tests.sh [pytest args]   (run on the GPU box through gpurun)
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 180 "$@" 2>&1 | tee gpurun_out/pytest_gpu.log | tail -40
