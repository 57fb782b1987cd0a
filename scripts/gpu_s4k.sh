cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_shard.py -x -q -m gpu --timeout 300 2>&1 | tail -4
SUBG_PROFILE_HOST=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --e2e-steps 1 > gpurun_out/s4k_bench2.json 2> gpurun_out/s4k_bench2.err
grep "exchange:" gpurun_out/s4k_bench2.err | tail -4
grep -o '"sharded".*' gpurun_out/s4k_bench2.json
