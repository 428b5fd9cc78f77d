# CUDA-core lin_out + warm rounds: sweep (toggles), sizes, whole suite with the option on, timeline and bench off / on
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sweep" 2>&1 | tail -8 | cut -c1-300
export DINER_TC_WARM_ROUNDS=1
for n in 8192 262144; do
  echo "== rays $n"; DINER_TC_DBG_SKIP=$([ $n = 8192 ] && echo 512 || echo 0) timeout 200 python tools/profile_run.py parity $n 1 2>&1 | grep "cta 0\|rep 0\|kernel cycles\|Error" | cut -c1-420
done
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | cut -c1-300
unset DINER_TC_WARM_ROUNDS
for e in 0 1 0 1; do
  DINER_TC_WARM_ROUNDS=$e timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2w_bench_err.txt | grep "^{" > gpurun_out/r2w_bench_warm$e.json
  python -c "
import json;d=json.load(open('gpurun_out/r2w_bench_warm$e.json'));print('warm',$e,d['value'],d['ms_per_step'],d['roofline']['frac'],d['clocks'],d.get('parity'))" || tail -3 gpurun_out/r2w_bench_err.txt
done
