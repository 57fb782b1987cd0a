cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
{ for st in 0 3 4 5; do SUBG_SAMPLER_STOP=$st python scripts/sampler_bench.py ppa 3; done; } 2>&1 | grep -v Warning | tee gpurun_out/s4h_sweep.txt
{ for st in 0 4 5; do SUBG_SAMPLER_STOP=$st python scripts/sampler_bench.py dblp 3; done; } 2>&1 | grep -v Warning | tee -a gpurun_out/s4h_sweep.txt
