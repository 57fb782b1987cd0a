cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_spjoin.py -x -q -m gpu --timeout 120 -k "empty or fused or bad" 2>&1 | tail -5
