cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -x -q -m gpu --timeout 600 2>&1 | tee gpurun_out/final_pytest.log | tail -4 ) 2>&1 | tail -7
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo rc=$?
python -c "
import json; d=json.load(open('gpurun_out/final_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['spjoin']['value'], d['spjoin']['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'])"
python scripts/walks_bench.py 2>/dev/null | tail -1 | tee gpurun_out/final_walks.json
