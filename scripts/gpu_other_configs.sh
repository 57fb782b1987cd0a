# bench lines of the other BASELINE configs (C1 collab with the reference timed beside it, C4 dblp)
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
python bench.py --workload collab --spjoin-batch 11264 > gpurun_out/bench_collab.json 2> gpurun_out/bench_collab.err; echo rc=$?
python bench.py --workload dblp --spjoin-batch 2048 --no-cpu-baseline > gpurun_out/bench_dblp.json 2> gpurun_out/bench_dblp.err; echo rc=$?
python -c "
import json
for w in ('collab','dblp'):
    d=json.load(open(f'gpurun_out/bench_{w}.json')); print(w, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['spjoin']['value'], d.get('cpu_baseline',{}).get('value'))"
