# usage (on the GPU box via gpurun): bash scripts/gpu_bench.sh [bench args]
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "rc=$?"; tail -3 gpurun_out/bench.err | cut -c1-300; cat gpurun_out/bench.json
