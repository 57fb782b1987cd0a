cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
N=${1:-4}
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 scripts/nccl_probe.py 2>&1 | grep "nccl probe"; }
run 29521 | tee gpurun_out/nccl_probe_$N.txt
NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32 run 29522 | tee -a gpurun_out/nccl_probe_$N.txt
NCCL_MIN_NCHANNELS=32 NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32 NCCL_P2P_NET_CHUNKSIZE=4194304 run 29523 | tee -a gpurun_out/nccl_probe_$N.txt
