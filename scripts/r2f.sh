# round 2, call f (1 GPU): ncu source-level capture of the hash sampler kernel on ppa
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gset_hash -s 1 -c 1 -f -o gpurun_out/r2f_hash python bench.py --workload ppa --quick --steps 1 --warmup 1 > gpurun_out/r2f_hash.log 2>&1
ls -la gpurun_out/r2f_hash.ncu-rep
