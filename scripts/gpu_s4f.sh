cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spjoin.py tests/test_gpu_ppr.py tests/test_gpu_fullsize.py -x -q -m gpu --timeout 300 2>&1 | tee gpurun_out/s4f_pytest.log | tail -5
python scripts/spjoin_probe.py ppa 21504 2>&1 | grep -v Warn | tee gpurun_out/s4f_probe.txt
python scripts/spjoin_probe.py ppa 1024 2>&1 | grep -v Warn | tee -a gpurun_out/s4f_probe.txt
python scripts/spjoin_probe.py collab 1024 2>&1 | grep -v Warn | tee -a gpurun_out/s4f_probe.txt
python scripts/c5_bench.py 2 > gpurun_out/s4f_c5.json 2> gpurun_out/s4f_c5.err; tail -5 gpurun_out/s4f_c5.err; cat gpurun_out/s4f_c5.json
