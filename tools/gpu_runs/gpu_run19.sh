cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -q --timeout 400 --tb=short -s -k "config3_size" 2>&1 | grep -E "config-3|passed|failed|FAILED|Error|assert" | cut -c1-250
timeout 900 python bench.py --workload train256 --steps 3 --warmup 1 > gpurun_out/r2q_bench_train256.json 2> gpurun_out/r2q_bench_train256.err; cat gpurun_out/r2q_bench_train256.json | cut -c1-1500; tail -2 gpurun_out/r2q_bench_train256.err
