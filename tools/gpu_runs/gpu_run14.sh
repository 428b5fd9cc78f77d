# warm rounds (default) + features computed before the K block 0 release: sweep, timeline, bench
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sweep or properties_and_edges or composite_stagewise" 2>&1 | tail -4 | cut -c1-300
DINER_TC_DBG_SKIP=512 timeout 200 python tools/profile_run.py parity 8192 1 2>&1 | grep "cta 0\|rep 0\|kernel cycles\|Error" | cut -c1-420
for e in 1 0 1; do
  DINER_TC_WARM_ROUNDS=$e timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2w_bench_err.txt | grep "^{" > gpurun_out/r2w_bench_warm$e.json
  python -c "
import json;d=json.load(open('gpurun_out/r2w_bench_warm$e.json'));print('warm',$e,d['value'],d['ms_per_step'],d['roofline']['frac'],d['clocks'])" || tail -3 gpurun_out/r2w_bench_err.txt
done
