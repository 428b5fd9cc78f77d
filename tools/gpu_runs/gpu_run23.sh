# launch list of one training step (config 3) with the final kernels
cd $GRAFT_REPO_ROOT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2z_train_launches.csv python bench.py --workload train256 --steps 1 --warmup 1 > gpurun_out/r2z_ncu_train.log 2>&1; tail -1 gpurun_out/r2z_ncu_train.log | cut -c1-300
