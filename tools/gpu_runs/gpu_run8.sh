cd $GRAFT_REPO_ROOT
for F in 0 1; do
DINER_TC_FUSED=$F DINER_TC_DBG_SKIP=512 timeout 300 python tools/profile_run.py parity 8192 2 2>&1 | grep -E "ts\]|rep" | cut -c1-400 | tee gpurun_out/r2h_ts_fused$F.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_pair_kernel -s 1 -c 1 -o gpurun_out/r2h_fused_parity python tools/profile_run.py parity 8192 2 > gpurun_out/r2h_ncu_full.log 2>&1; tail -3 gpurun_out/r2h_ncu_full.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2h_ncu_list.log 2>&1; tail -2 gpurun_out/r2h_ncu_list.log | cut -c1-200
