cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
{
for h in 0 1 2 3; do python scripts/sampler_bench.py ppa 5; done
RANKS=1 python scripts/sampler_bench.py ppa 5
python scripts/sampler_bench.py collab 5
python scripts/sampler_bench.py dblp 5
} 2>&1 | grep -v Warning | tee gpurun_out/sweep1.txt
ncu --set full --clock-control none --import-source on -k regex:gset_sample -s 2 -c 1 -f -o gpurun_out/r1c_sampler python scripts/sampler_bench.py ppa 1 > gpurun_out/r1c_sampler.log 2>&1
ls -la gpurun_out | tail -5
