# round 2, call p (8 GPUs): the exchange at the twitter size -- per-rank pull times, peer vs NCCL staging, grid size
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
run() { tag=$1; shift
  env BENCH_PER_RANK=1 "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --workload twitter --steps 2 --warmup 2 --no-replicas --quick > gpurun_out/r2p_$tag.json 2> gpurun_out/r2p_$tag.err
  echo "== $tag rc=$?"; grep "\[bench\] rank\|parity" gpurun_out/r2p_$tag.err | cut -c1-200 | sort | head -20
  python - gpurun_out/r2p_$tag.json <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); s=d["sharded"]
    print({k:v for k,v in s.items() if k in ("mode","ms_per_pass","sampler_kernel_ms","exchange_ms","pull_kernel_ms","pull_GBps_per_gpu","parity_ok","parity_error")})
    print("nccl:", {k:v for k,v in (s.get("nccl_staged") or {}).items() if k in ("ms_per_pass","exchange_ms","pull_kernel_ms","exchange_GBps_per_gpu")})
except Exception as e: print("no json", e)
P
}
run default X=1
run blocks2 SUBG_XCHG_BLOCKS=296
