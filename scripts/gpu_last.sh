cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
python bench.py --scale 0.05 --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline 2>gpurun_out/tiny.err | cut -c1-400; tail -2 gpurun_out/tiny.err | cut -c1-200
