# usage (GPU box): bash scripts/gpu_ncu.sh <tag> [bench args]  -> gpurun_out/<tag>_launches.csv, <tag>_sampler.ncu-rep, <tag>_spjoin.ncu-rep
cd ${GRAFT_REPO_ROOT:-.}
tag=$1; shift
mkdir -p gpurun_out
BA="--steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline $@"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py $BA > gpurun_out/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gset_sample -s 1 -c 1 -f -o gpurun_out/${tag}_sampler python bench.py $BA > gpurun_out/${tag}_sampler.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spjoin_kernel -s 3 -c 1 -f -o gpurun_out/${tag}_spjoin python bench.py $BA > gpurun_out/${tag}_spjoin.log 2>&1
ls -la gpurun_out/
