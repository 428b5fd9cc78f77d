cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -q --timeout 300 --tb=short -s -k "backward or training" 2>&1 | grep -E "backward \(|passed|failed|FAILED|Error" | cut -c1-250
timeout 900 python bench.py --workload train256 --steps 3 --warmup 1 > gpurun_out/r2p_bench_train256.json 2> gpurun_out/r2p_bench_train256.err; cut -c1-330 gpurun_out/r2p_bench_train256.json; tail -2 gpurun_out/r2p_bench_train256.err
