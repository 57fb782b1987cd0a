# round 2, call 3m (2 GPUs): linked shards over real peer mappings (worker test) + the bench line with its linked block
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r3m_ppa2.json 2> gpurun_out/r3m_ppa2.err
echo "ppa2 rc=$?"; grep -v "^\[W\|NCCL\|^$\|^\*\*\*\|OMP_NUM" gpurun_out/r3m_ppa2.err | tail -4 | cut -c1-300
python - <<'P'
import json
d=json.loads(open("gpurun_out/r3m_ppa2.json").read().strip().splitlines()[-1]); s=d["sharded"]
print("value %.4g ms %.3f"%(d["value"], d["ms_per_step"]), {k:s.get(k) for k in ("mode","ms_per_pass","sampler_kernel_ms","exchange_ms","pull_kernel_ms","pull_GBps_per_gpu","parity_ok")}, "e2e", d["e2e"]["value"], "replicas", (d.get("replicas") or {}).get("value"), "gather", d["roofline"].get("gather",{}).get("frac_of_gather_rate"))
for b in d.get("spjoin_batches") or []: print("spjoin", b.get("batch"), b.get("value"), (b.get("stream") or {}).get("value"), b.get("error"))
P
timeout 600 python -m pytest tests/test_gpu_shard.py -x -q -m gpu --timeout 500 2>&1 | tail -2
python - <<'P'
import json
d=json.loads(open("gpurun_out/r3m_ppa2.json").read().strip().splitlines()[-1]); print("linked:", json.dumps(d.get("linked"))[:1500])
P
