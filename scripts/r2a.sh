# round 2, call a (1 GPU): all gpu tests + default bench
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tee gpurun_out/r2a_pytest.log | tail -15
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2a_bench.err | cut -c1-300; cut -c1-1500 gpurun_out/r2a_bench.json
