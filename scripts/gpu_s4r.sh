cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 900 python examples/link_prediction.py --steps 400 2>&1 | grep -v "Warning\|Start sampling" | tee gpurun_out/s4r_linkpred_mean.txt | tail -4
timeout 900 python examples/link_prediction.py --steps 400 --aggr attn --model-seeds 2 --sample-seeds 3 2>&1 | grep -v "Warning\|Start sampling" | tee gpurun_out/s4r_linkpred_attn.txt | tail -3
# 4-GPU check of the scaling bench rides along in the next call
