# round 2, call x (8 GPUs): final lines -- ppa sharded over 8 GPUs (BASELINE configs[1]), twitter shape (configs[4]) with the
# default exchange policy (windowed pull, 2 blocks per SM above 4 GB) and the windowed pull at 8 blocks per SM for comparison
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -12 > gpurun_out/r2x_topo.txt
show() { python - "$1" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.4g ms/step %.3f scaling %s" % (d["value"], d["ms_per_step"], d["scaling"]))
    s=d.get("sharded") or {}; print({k:v for k,v in s.items() if k not in ("what","parity_check","what_exchange_ms_covers","nccl_staged","nvlink")})
    print("nccl:", {k:v for k,v in (s.get("nccl_staged") or {}).items() if k in ("mode","ms_per_pass","exchange_ms","pull_kernel_ms","exchange_GBps_per_gpu")})
    print("replicas:", d.get("replicas")); print("e2e:", d["e2e"])
    for b in d.get("spjoin_batches") or []: print("spjoin", b.get("batch"), b.get("value"), (b.get("stream") or {}).get("value"), b.get("error"))
except Exception as e: print("no json", e)
P
}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2x_ppa8.json 2> gpurun_out/r2x_ppa8.err
echo "ppa8 rc=$?"; grep -v "^\[W\|NCCL\|^$\|^\*\*\*\|OMP_NUM" gpurun_out/r2x_ppa8.err | tail -4 | cut -c1-300; show gpurun_out/r2x_ppa8.json
BENCH_PER_RANK=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --workload twitter --steps 3 --warmup 2 --no-replicas > gpurun_out/r2x_twitter8.json 2> gpurun_out/r2x_twitter8.err
echo "twitter8 rc=$?"; grep "\[bench\] rank\|parity" gpurun_out/r2x_twitter8.err | cut -c1-200 | sort | head -20; show gpurun_out/r2x_twitter8.json
BENCH_PER_RANK=1 SUBG_XCHG_BLOCKS=1184 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --workload twitter --steps 2 --warmup 2 --no-replicas --no-exchange-compare --quick > gpurun_out/r2x_twitter8_ticket1184.json 2> gpurun_out/r2x_twitter8_ticket1184.err
echo "twitter8 windowed 1184 rc=$?"; grep "\[bench\] rank" gpurun_out/r2x_twitter8_ticket1184.err | cut -c1-200 | sort | head -10; show gpurun_out/r2x_twitter8_ticket1184.json
