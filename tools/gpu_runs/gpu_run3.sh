cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 -k "sweep or properties_and_edges" 2>&1 | tail -5
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -s 2>&1 | grep -E "(NV=|[a-z0-9_]+(/[a-z0-9]+)?: |full-size|whole-image|gen_rays|depth2normal|fast mode|training|softplus|[0-9]+ (passed|failed)|FAILED|Error)" | cut -c1-330 > gpurun_out/r2c_pytest_prints.log
cat gpurun_out/r2c_pytest_prints.log
for cfg in "3 0" "3 16" "4 0" "2 0"; do
  set -- $cfg
  DINER_TC_TAIL_KB=$1 DINER_TC_DBG_SKIP=$2 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_tail$1_skip$2.json 2> gpurun_out/r2c_bench_tail$1_skip$2.err
  python -c "
import json;d=json.load(open('gpurun_out/r2c_bench_tail$1_skip$2.json'));print('tail',$1,'dbg_skip',$2,d['value'],d['roofline']['frac'],d['roofline']['stage_ms_per_step'],d['clocks'])"
done
DINER_TC_DBG_SKIP=512 DINER_TC_TAIL_KB=3 timeout 300 python tools/profile_run.py parity 8192 2 2>&1 | grep -E "ts\]|rep" | tee gpurun_out/r2c_ts_tail3.log
