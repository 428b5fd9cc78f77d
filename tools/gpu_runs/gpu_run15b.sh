# helpers stage K block 0 of warm tiles (dbg_skip 1024 = off); POST x_c load in halves (2048 = off): sweep, then A/B bench
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sweep or properties_and_edges or composite_stagewise" 2>&1 | tail -4 | cut -c1-300
for e in 0 3072 1024 2048 0 3072; do
  DINER_TC_DBG_SKIP=$e timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2x_bench_err.txt | grep "^{" > gpurun_out/r2x_bench_skip$e.json
  python -c "
import json;d=json.load(open('gpurun_out/r2x_bench_skip$e.json'));print('dbg_skip',$e,d['value'],d['ms_per_step'],d['roofline']['frac'],d['clocks'])" || tail -3 gpurun_out/r2x_bench_err.txt
done
