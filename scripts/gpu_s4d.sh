cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
free -g | head -2
timeout 600 python -m pytest tests/test_gpu_ingest.py -x -q -m gpu --timeout 300 2>&1 | tee gpurun_out/s4d_ingest.log | tail -15
timeout 900 python -m pytest tests/test_gpu_c5.py -x -q -s -m gpu --timeout 800 2>&1 | tee gpurun_out/s4d_c5.log | tail -25
