cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -x -q -m gpu --timeout 600 2>&1 | tee gpurun_out/s4m_pytest.log | tail -5 ) 2>&1 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/s4m_bench.json 2> gpurun_out/s4m_bench.err; echo rc=$?
cat gpurun_out/s4m_bench.json
( time python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s4m_ref.json 2> gpurun_out/s4m_ref.err ) 2>&1 | tail -3; cat gpurun_out/s4m_ref.json
