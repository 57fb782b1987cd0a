cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
bash scripts/gpu_ncu.sh r1m > gpurun_out/r1m_ncu.log 2>&1
tail -3 gpurun_out/r1m_ncu.log
ls -la gpurun_out/*.ncu-rep
