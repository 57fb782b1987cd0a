cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_walks.py -x -q -m gpu --timeout 300 2>&1 | tail -4
timeout 600 python examples/link_prediction.py --steps 400 2>&1 | grep -v "Warning\|Start sampling" | tee gpurun_out/s4p_linkpred_mean.txt | tail -9
timeout 600 python examples/link_prediction.py --steps 400 --aggr attn --model-seeds 2 2>&1 | grep -v "Warning\|Start sampling" | tee gpurun_out/s4p_linkpred_attn.txt | tail -7
