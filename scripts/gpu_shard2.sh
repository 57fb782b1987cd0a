cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_shard.py -x -q -m gpu --timeout 300 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/shard_bench$N.json 2> gpurun_out/shard_bench$N.err
echo rc=$?; grep -o '"sharded".*' gpurun_out/shard_bench$N.json
SUBG_PROFILE_HOST=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline 2>&1 >/dev/null | grep "exchange:" | tail -3
