# round 2, call t (2 GPUs): the NCCL / peer exchange tests that need two devices, then the pull schedule at the twitter size
# (11 GB pulled per GPU from one peer): grid-stride vs windowed (SUBG_XCHG_TICKET), blocks per GPU
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_shard.py -x -q -m gpu --timeout 500 2>&1 | tee gpurun_out/r2t_pytest_shard.log | tail -4
SUBG_XCHG_TICKET=1 timeout 600 python -m pytest tests/test_gpu_shard.py -x -q -m gpu --timeout 500 2>&1 | tee gpurun_out/r2t_pytest_shard_ticket.log | tail -4
run() { tag=$1; shift
  env BENCH_PER_RANK=1 "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --workload twitter --steps 2 --warmup 2 --no-replicas --no-exchange-compare --quick > gpurun_out/r2t_$tag.json 2> gpurun_out/r2t_$tag.err
  echo "== $tag rc=$?"; grep "\[bench\] rank\|parity" gpurun_out/r2t_$tag.err | cut -c1-200 | sort | head -8
  python - gpurun_out/r2t_$tag.json <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); s=d["sharded"]
    print({k:v for k,v in s.items() if k in ("mode","ms_per_pass","sampler_kernel_ms","exchange_ms","pull_kernel_ms","pull_GBps_per_gpu","parity_ok","parity_error")})
except Exception as e: print("no json", e)
P
}
run stride_1184 SUBG_XCHG_TICKET=0
run stride_296 SUBG_XCHG_TICKET=0 SUBG_XCHG_BLOCKS=296
run ticket_1184 SUBG_XCHG_TICKET=1
run ticket_592 SUBG_XCHG_TICKET=1 SUBG_XCHG_BLOCKS=592
run ticket_296 SUBG_XCHG_TICKET=1 SUBG_XCHG_BLOCKS=296
