# early_lin after the late-waiter fix: sizes, timeline of a warm tile with the option off / on, bench off / on
cd $GRAFT_REPO_ROOT
for e in 0 1; do
  echo "== early_lin $e timeline"
  DINER_TC_EARLY_LIN=$e DINER_TC_DBG_SKIP=512 timeout 300 python tools/profile_run.py parity 8192 1 2>&1 | grep "cta 0\|rep 0\|kernel cycles" | cut -c1-420
done
export DINER_TC_EARLY_LIN=1
for n in 131072 262144; do
  echo "== rays $n"; timeout 200 python tools/profile_run.py parity $n 1 2>&1 | grep -v "^\[ts\]" | tail -3 | cut -c1-600
done
for e in 0 1 0 1; do
  DINER_TC_EARLY_LIN=$e timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2v_bench_err.txt | grep "^{" > gpurun_out/r2v_bench_early$e.json
  python -c "
import json;d=json.load(open('gpurun_out/r2v_bench_early$e.json'));print('early',$e,d['value'],d['ms_per_step'],d['roofline']['frac'],d['clocks'])" || tail -3 gpurun_out/r2v_bench_err.txt
done
