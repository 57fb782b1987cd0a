# round 2, call b (2 GPUs): exchange over peer memory between two processes + new parity tests + 2-GPU bench
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -12 > gpurun_out/r2b_topo.txt
timeout 900 python -m pytest tests/test_gpu_shard.py tests/test_gpu_philox_parity.py -x -q -m gpu --timeout 600 2>&1 | tee gpurun_out/r2b_pytest.log | tail -25
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2b_bench2.json 2> gpurun_out/r2b_bench2.err
echo "bench rc=$?"; grep -v "^\[W\|NCCL\|^$" gpurun_out/r2b_bench2.err | tail -8 | cut -c1-400; python - <<'P'
import json
try:
    d=json.loads(open("gpurun_out/r2b_bench2.json").read().strip().splitlines()[-1]); print(json.dumps(d.get("sharded"))); print(d["value"], d["ms_per_step"])
except Exception as e: print("no json", e)
P
