# round 2, call e (1 GPU): where does the hash kernel spend its time (phase stops)
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
for st in 6 7 1 2 0; do
  SUBG_SAMPLER_STOP=$st timeout 300 python bench.py --workload ppa --quick --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('ppa hash stop=$st kernel_ms %.3f' % r['kernel_ms_per_launch'])"
done
SUBG_SAMPLER_HASH=0 SUBG_SAMPLER_STOP=1 timeout 300 python bench.py --workload ppa --quick --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('ppa sort-kernel stop=1 kernel_ms %.3f' % r['kernel_ms_per_launch'])"
for st in 6 7 1 2 0; do
  SUBG_SAMPLER_STOP=$st timeout 300 python bench.py --workload collab --quick --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('collab hash stop=$st kernel_ms %.3f' % r['kernel_ms_per_launch'])"
done
