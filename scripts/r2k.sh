# round 2, call k (2 GPUs): sharded tests + the N=2 bench line (sharded headline, parity check, peer vs NCCL staging)
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shard.py -x -q -m gpu --timeout 600 2>&1 | tee gpurun_out/r2k_pytest.log | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2k_bench2.json 2> gpurun_out/r2k_bench2.err
echo "bench rc=$?"; grep -v "^\[W\|NCCL\|^$\|^\*\*\*\|OMP_NUM" gpurun_out/r2k_bench2.err | tail -8 | cut -c1-400
python - <<'P'
import json
try:
    d=json.loads(open("gpurun_out/r2k_bench2.json").read().strip().splitlines()[-1])
    print("value %.4g ms/step %.3f scaling %s" % (d["value"], d["ms_per_step"], d["scaling"]))
    s=d["sharded"]; print({k:v for k,v in s.items() if k not in ("what","parity_check","what_exchange_ms_covers","nccl_staged","nvlink")})
    print("nccl:", {k:v for k,v in (s.get("nccl_staged") or {}).items() if k in ("mode","ms_per_pass","exchange_ms","pull_kernel_ms","exchange_GBps_per_gpu")})
    print("replicas:", d.get("replicas")); print("e2e:", d["e2e"])
    for b in d["spjoin_batches"]: print("spjoin", b.get("batch"), b.get("value"), (b.get("stream") or {}).get("value"))
except Exception as e: print("no json", e)
P
