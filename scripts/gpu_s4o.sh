cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 600 python examples/link_prediction.py --steps 300 2>&1 | grep -v Warning | tee gpurun_out/s4o_linkpred_mean.txt | tail -12
timeout 600 python examples/link_prediction.py --steps 300 --aggr attn --model-seeds 2 2>&1 | grep -v Warning | tee gpurun_out/s4o_linkpred_attn.txt | tail -8
