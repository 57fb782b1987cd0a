set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_gset.py -x -q -m gpu 2>&1 | tail -30
