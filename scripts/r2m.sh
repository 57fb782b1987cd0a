# round 2, call m (2 GPUs): why is the exchange slow at the twitter size?  (a quarter-size twitter graph on 2 GPUs has the
# same per-rank shard as the full graph on 8)
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
SUBG_PROFILE_HOST=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload twitter --scale 0.25 --steps 2 --warmup 1 --no-replicas --quick > gpurun_out/r2m_tw.json 2> gpurun_out/r2m_tw.err
echo rc=$?; grep "subg host ms" gpurun_out/r2m_tw.err | grep -v "setup=" | head -16 | cut -c1-400
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2m_tw.json").read().strip().splitlines()[-1]); s=d["sharded"]
print({k:v for k,v in s.items() if k in ("mode","ms_per_pass","sampler_kernel_ms","exchange_ms","pull_kernel_ms","pull_GBps_per_gpu","received_bytes_per_gpu","parity_ok")})
print("nccl", {k:v for k,v in (s.get("nccl_staged") or {}).items() if k in ("ms_per_pass","exchange_ms","pull_kernel_ms")})
P
