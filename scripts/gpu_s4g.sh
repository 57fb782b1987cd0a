cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spjoin.py tests/test_gpu_ppr.py -x -q -m gpu --timeout 300 2>&1 | tee gpurun_out/s4g_pytest.log | tail -5
python scripts/spjoin_probe.py ppa 21504 2>&1 | grep -v Warn | tee gpurun_out/s4g_probe.txt
python scripts/spjoin_probe.py ppa 1024 2>&1 | grep -v Warn | tee -a gpurun_out/s4g_probe.txt
python scripts/spjoin_probe.py dblp 2048 2>&1 | grep -v Warn | tee -a gpurun_out/s4g_probe.txt
