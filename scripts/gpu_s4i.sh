# 2-GPU: NCCL shard test + bench with the sharded (strong scaling) block
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -3
timeout 600 python -m pytest tests/test_gpu_shard.py -x -q -m gpu --timeout 300 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s4i_bench2.json 2> gpurun_out/s4i_bench2.err
echo rc=$?; tail -3 gpurun_out/s4i_bench2.err; cat gpurun_out/s4i_bench2.json
