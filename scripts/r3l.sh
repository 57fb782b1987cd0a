# round 2, call 3l (1 GPU): linked shards (subg_xchg_stage / subg_xchg_link) in one process + the exchange tests
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_shard.py -x -q -m gpu --timeout 500 2>&1 | tee gpurun_out/r3l_pytest.log | tail -15
