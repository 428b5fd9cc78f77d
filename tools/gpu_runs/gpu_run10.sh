# early_lin validation: gate tests (the sweep toggles the option in both launch modes), then the whole parity file with the option on, bench both
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sweep" 2>&1 | tail -5
export DINER_TC_EARLY_LIN=1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
unset DINER_TC_EARLY_LIN
for e in 0 1 0 1; do
  DINER_TC_EARLY_LIN=$e timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | grep "^{" > gpurun_out/r2v_bench_early$e.json
  python -c "
import json;d=json.load(open('gpurun_out/r2v_bench_early$e.json'));print('early',$e,d['value'],d['ms_per_step'],d['roofline']['frac'],d['clocks'])"
done
DINER_TC_EARLY_LIN=1 DINER_TC_DBG_SKIP=512 timeout 300 python tools/profile_run.py parity 8192 1 2>&1 | tail -40 > gpurun_out/r2v_timeline_early1.txt
