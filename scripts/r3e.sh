# round 2, call 3e (1 GPU): acceptance "Hits@50 unchanged" with LSTM pooling (model.py:63-66) and the compiled reference run
# the way users run it (nthread = -1) as a fourth source of walks
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 900 python -u tests/acceptance_hits50.py --aggr lstm --steps 300 --model-seeds 2 --sample-seeds 3 > gpurun_out/r3e_hits50_lstm.txt 2> gpurun_out/r3e_hits50_lstm.err
echo "rc=$?"; tail -12 gpurun_out/r3e_hits50_lstm.txt; tail -3 gpurun_out/r3e_hits50_lstm.err
