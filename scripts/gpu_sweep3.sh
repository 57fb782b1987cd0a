cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
{
for st in 1 2 3 0; do SUBG_SAMPLER_STOP=$st python scripts/sampler_bench.py ppa 3; done
for st in 1 2 3 0; do SUBG_SAMPLER_STOP=$st SUBG_SAMPLER_HINTS=0 python scripts/sampler_bench.py ppa 3; done
for st in 1 2 3 0; do SUBG_SAMPLER_STOP=$st python scripts/sampler_bench.py dblp 3; done
} 2>&1 | grep -v Warning | tee gpurun_out/sweep3.txt
ncu --set full --clock-control none --import-source on -k regex:gset_sample -s 2 -c 1 -f -o gpurun_out/r1d_sampler python scripts/sampler_bench.py ppa 1 > gpurun_out/r1d_sampler.log 2>&1
