cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu --timeout 600 2>&1 | tail -15
bash scripts/gpu_ncu.sh r1g > gpurun_out/r1g_ncu.log 2>&1
tail -3 gpurun_out/r1g_ncu.log
{ for st in 1 2 3 0; do SUBG_SAMPLER_STOP=$st python scripts/sampler_bench.py dblp 3; done; } 2>&1 | grep -v Warning | tee gpurun_out/sweep_dblp.txt
