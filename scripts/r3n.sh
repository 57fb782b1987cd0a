# round 2, call 3n (1 GPU): the whole gpu suite + smoke on the final code of the round
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 900 2>&1 | tee gpurun_out/r3n_pytest.log | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/r3n_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('value %.4g ms/step %.3f kernel %.3f frac %.4f e2e %.4g scaling %s launches %s'%(d['value'], d['ms_per_step'], r['kernel_ms_per_launch'], r['frac'], d['e2e']['value'], d['scaling'], d['gpu_launches']))"
