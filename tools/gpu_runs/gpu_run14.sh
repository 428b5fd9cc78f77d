cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -q --timeout 300 --tb=short -k "backward or training" 2>&1 | tail -3
DINER_B200_BACKWARD_TC=1 timeout 900 python bench.py --workload train256 --steps 3 --warmup 1 > gpurun_out/r2m_bench_train256_tc1.json 2> gpurun_out/r2m_bench_train256_tc1.err; cut -c1-700 gpurun_out/r2m_bench_train256_tc1.json; tail -2 gpurun_out/r2m_bench_train256_tc1.err
