# final pass of the round: whole GPU suite, smoke, default bench line (with cpu_baseline / parity block), fast mode, two-kernel path,
# timeline, ncu launch list of a bench step, ncu --set full of the fused kernel, reference arm
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | cut -c1-300
timeout 900 python bench.py 2>gpurun_out/r2z_bench_err.txt | grep "^{" > gpurun_out/r2z_bench_default.json; python -c "
import json;d=json.load(open('gpurun_out/r2z_bench_default.json'));print(d['value'],d['e2e'],d['ms_per_step'],d['roofline'],d['clocks'],d.get('parity'),d['cpu_baseline'],d['gpu_launches'])" || tail -3 gpurun_out/r2z_bench_err.txt
DINER_B200_MODE=fast timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | grep "^{" > gpurun_out/r2z_bench_fast.json; python -c "
import json;d=json.load(open('gpurun_out/r2z_bench_fast.json'));print('fast',d['value'],d['roofline']['frac'])"
DINER_TC_FUSED=0 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | grep "^{" > gpurun_out/r2z_bench_two_kernel.json; python -c "
import json;d=json.load(open('gpurun_out/r2z_bench_two_kernel.json'));print('two-kernel',d['value'],d['roofline']['frac'])"
DINER_TC_DBG_SKIP=512 timeout 300 python tools/profile_run.py parity 8192 1 2>&1 | grep -E "ts\]|rep" | cut -c1-420 > gpurun_out/r2z_timeline.txt; head -3 gpurun_out/r2z_timeline.txt | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2z_ncu_list.log 2>&1; tail -1 gpurun_out/r2z_ncu_list.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_pair_kernel -s 1 -c 1 -o gpurun_out/r2z_fused_parity python tools/profile_run.py parity 8192 2 > gpurun_out/r2z_ncu_full.log 2>&1; tail -2 gpurun_out/r2z_ncu_full.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | grep "^{" | cut -c1-400
