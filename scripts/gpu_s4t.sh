cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 1200 python examples/link_prediction.py --steps 400 2>&1 | grep -v "Warning\|Start sampling" | tee gpurun_out/s4t_linkpred_mean.txt | tail -8
