cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 -k "sweep or properties_and_edges or encode_path or full_size" 2>&1 | tail -4
for cfg in "3 0" "3 16" "4 0" "4 16" "2 16"; do
  set -- $cfg
  DINER_TC_TAIL_KB=$1 DINER_TC_DBG_SKIP=$2 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench_tail$1_skip$2.json 2> gpurun_out/r2d_bench_tail$1_skip$2.err
  python -c "
import json;d=json.load(open('gpurun_out/r2d_bench_tail$1_skip$2.json'));print('tail',$1,'dbg_skip',$2,d['value'],d['roofline']['frac'],d['roofline']['stage_ms_per_step'],d['clocks'])"
done
DINER_TC_DBG_SKIP=528 DINER_TC_TAIL_KB=3 timeout 300 python tools/profile_run.py parity 8192 2 2>&1 | grep -E "ts\]|rep" | tee gpurun_out/r2d_ts_tail3.log
DINER_TC_DBG_SKIP=528 DINER_TC_TAIL_KB=4 timeout 300 python tools/profile_run.py parity 8192 2 2>&1 | grep -E "ts\]|rep" | tee gpurun_out/r2d_ts_tail4.log
