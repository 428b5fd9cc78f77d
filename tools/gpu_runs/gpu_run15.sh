cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -q --timeout 300 --tb=short -k "backward or training" 2>&1 | tail -3
timeout 900 python bench.py --workload train256 --steps 3 --warmup 1 > gpurun_out/r2n_bench_train256.json 2> gpurun_out/r2n_bench_train256.err; cut -c1-400 gpurun_out/r2n_bench_train256.json; tail -2 gpurun_out/r2n_bench_train256.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2n_train_launches.csv python bench.py --workload train256 --steps 1 --warmup 1 > gpurun_out/r2n_ncu_train.log 2>&1; tail -2 gpurun_out/r2n_ncu_train.log | cut -c1-200
