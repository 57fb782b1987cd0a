# round 2, call 3c (1 GPU): staged upload of the pageable CSR -- threads / chunk size
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
nproc; lscpu | grep -E "Model name|Socket|NUMA node\(s\)|^CPU\(s\)" 
PROBE_BASE=1 python scripts/h2d_probe.py
for t in 4 8 12 16; do SUBG_COPY_THREADS=$t python scripts/h2d_probe.py; done
for c in 2 4 16; do SUBG_COPY_THREADS=8 SUBG_COPY_CHUNK_MB=$c python scripts/h2d_probe.py; done
