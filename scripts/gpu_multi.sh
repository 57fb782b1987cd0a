cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
N=${1:-4}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/multi_bench$N.json 2> gpurun_out/multi_bench$N.err
echo rc=$?; tail -2 gpurun_out/multi_bench$N.err | cut -c1-300; cat gpurun_out/multi_bench$N.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 1 --warmup 0 --ref-seconds 3 2>/dev/null | tail -1 | cut -c1-400
