cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --tb=short -k "sweep or composite_stagewise or properties_and_edges" 2>&1 | tail -3
DINER_TC_DBG_SKIP=512 timeout 300 python tools/profile_run.py parity 8192 1 2>&1 | grep -E "ts\]|rep" | cut -c1-420
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r2r_bench.json'));print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['roofline']['stage_ms_per_step'],d['clocks'])"
