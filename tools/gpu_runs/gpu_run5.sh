cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 -k "sweep or properties_and_edges or encode_path or cam_sweep" 2>&1 | tail -4
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8
timeout 600 python bench.py > gpurun_out/r2e_bench_default.json 2> gpurun_out/r2e_bench_default.err; tail -c 1500 gpurun_out/r2e_bench_default.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_pair_kernel -s 3 -c 2 -o gpurun_out/r2e_pair_parity python tools/profile_run.py parity 8192 2 > gpurun_out/r2e_ncu_full.log 2>&1; tail -3 gpurun_out/r2e_ncu_full.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2e_ncu_list.log 2>&1; tail -2 gpurun_out/r2e_ncu_list.log | cut -c1-300
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-600
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
