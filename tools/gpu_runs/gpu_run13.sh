cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -q --timeout 300 --tb=short -s -k "backward or training" 2>&1 | grep -E "backward \(|passed|failed|FAILED|Error|assert|training losses" | cut -c1-300
for B in 1 0; do
DINER_B200_BACKWARD_TC=$B timeout 900 python bench.py --workload train256 --steps 3 --warmup 1 > gpurun_out/r2l_bench_train256_tc$B.json 2> gpurun_out/r2l_bench_train256_tc$B.err; cut -c1-700 gpurun_out/r2l_bench_train256_tc$B.json; tail -2 gpurun_out/r2l_bench_train256_tc$B.err
done
