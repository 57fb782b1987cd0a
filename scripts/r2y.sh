# round 2, call y (4 GPUs): ppa sharded over 4 GPUs (the scaling table's missing point)
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2y_ppa4.json 2> gpurun_out/r2y_ppa4.err
echo "ppa4 rc=$?"; grep -v "^\[W\|NCCL\|^$\|^\*\*\*\|OMP_NUM" gpurun_out/r2y_ppa4.err | tail -4 | cut -c1-300
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2y_ppa4.json").read().strip().splitlines()[-1]); s=d["sharded"]
print("value %.4g ms %.3f"%(d["value"], d["ms_per_step"]), {k:s.get(k) for k in ("mode","ms_per_pass","sampler_kernel_ms","exchange_ms","pull_kernel_ms","pull_GBps_per_gpu","parity_ok")}, "e2e", d["e2e"]["value"], "replicas", (d.get("replicas") or {}).get("value"))
P
