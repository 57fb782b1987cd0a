# round 2, call c (1 GPU): full gpu test suite + smoke of every bench workload at reduced scale + default bench
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 600 2>&1 | tee gpurun_out/r2c_pytest.log | tail -8
for w in "collab 1.0" "dblp 0.1" "citation2-ppr 0.03" "twitter 0.003" "ppa 0.05"; do
  set -- $w
  timeout 300 python bench.py --workload $1 --scale $2 --steps 2 --warmup 1 --ref-seconds 2 > gpurun_out/r2c_smoke_$1.json 2> gpurun_out/r2c_smoke_$1.err
  echo "smoke $1 rc=$? $(cut -c1-200 gpurun_out/r2c_smoke_$1.json)"; grep -i "failed\|error\|Traceback" gpurun_out/r2c_smoke_$1.err | head -5
done
timeout 200 python bench.py --impl reference --workload ppa --scale 0.05 --steps 1 --warmup 1 --ref-seconds 2 2> gpurun_out/r2c_ref.err | cut -c1-300
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2c_bench.err | cut -c1-300; cut -c1-600 gpurun_out/r2c_bench.json
