cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -6
timeout 600 python bench.py > gpurun_out/r2j_bench_default.json 2> gpurun_out/r2j_bench_default.err; python -c "
import json;d=json.load(open('gpurun_out/r2j_bench_default.json'));print(d['value'],d['e2e'],d['roofline']['frac'],d['roofline']['stage_ms_per_step'],d['clocks'],d['gpu_launches'],d['parity'],d['cpu_baseline']['value'],d['gpu_eager_baseline'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_pair_kernel -s 1 -c 1 -o gpurun_out/r2j_fused_parity python tools/profile_run.py parity 8192 2 > gpurun_out/r2j_ncu_full.log 2>&1; tail -2 gpurun_out/r2j_ncu_full.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2j_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2j_ncu_list.log 2>&1
timeout 900 python bench.py --workload train256 --steps 2 --warmup 1 > gpurun_out/r2j_bench_train256.json 2> gpurun_out/r2j_bench_train256.err; cat gpurun_out/r2j_bench_train256.json | cut -c1-900; tail -3 gpurun_out/r2j_bench_train256.err
timeout 300 python bench.py --mode fast --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_bench_fast.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/r2j_bench_fast.json'));print('fast',d['value'],d['roofline']['frac'],d['parity'])"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
