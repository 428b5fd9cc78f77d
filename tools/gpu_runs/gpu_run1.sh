set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 -k "sweep or properties_and_edges" 2>&1 | tail -15 > gpurun_out/r2a_gate.log
cat gpurun_out/r2a_gate.log | tail -5
if grep -q "failed\|error" gpurun_out/r2a_gate.log; then echo GATE_FAILED; fi
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -s 2>&1 | grep -v "^$" | tail -150 > gpurun_out/r2a_pytest.log
tail -40 gpurun_out/r2a_pytest.log
for T in 0 2 3 4; do
  DINER_TC_TAIL_KB=$T timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_tail$T.json 2> gpurun_out/r2a_bench_tail$T.err
  python -c "
import json;d=json.load(open('gpurun_out/r2a_bench_tail$T.json'));print('tail',$T,d['value'],d['roofline']['frac'],d['roofline']['stage_ms_per_step'],d.get('parity'))"
done
