# round 2, call n (1 GPU): full gpu suite with the block cache + default bench line
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu --timeout 900 2>&1 | tee gpurun_out/r2n_pytest.log | tail -8
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2n_bench.err | cut -c1-300
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2n_bench.json").read().strip().splitlines()[-1])
print("value %.4g ms/step %.3f e2e %.4g (%.1f ms) e2e_numpy %.4g frac %.3f kms %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e_numpy"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"]))
for b in d["spjoin_batches"]:
    st=b.get("stream") or {}
    print("spjoin B", b.get("batch"), "gather %.3g q/s %.4f ms kshare %.2f | stream %s q/s ms %s host_us %s kshare %s" % (b.get("value"), b.get("ms_per_batch"), b["roofline"]["kernel_share_of_batch"], st.get("value"), st.get("ms_per_batch"), st.get("host_us_per_submit"), st.get("kernel_share_of_batch")))
print(d.get("cpu_baseline"))
P
