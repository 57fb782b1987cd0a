# round 2, call 3k (1 GPU): gpu suite + default bench line after the 64-register bound of the 19-keys-per-lane sampler
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 900 2>&1 | tee gpurun_out/r3k_pytest.log | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r3k_bench_ppa.json 2> gpurun_out/r3k_bench_ppa.err; echo "ppa rc=$?"
python - <<'P'
import json
d=json.loads(open("gpurun_out/r3k_bench_ppa.json").read().strip().splitlines()[-1]); r=d["roofline"]
print("value %.4g ms/step %.3f | kernel ms %.3f frac %.4f e2e %s cpu %s" % (d["value"], d["ms_per_step"], r["kernel_ms_per_launch"], r["frac"], d["e2e"]["value"], d["cpu_baseline"]["value"]))
P
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gset_sample -s 1 -c 1 -f -o gpurun_out/r3k_sampler python bench.py --steps 2 --warmup 1 --quick > gpurun_out/r3k_sampler.log 2>&1
