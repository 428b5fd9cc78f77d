cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --tb=short -s -k "sweep or cam_sweep or properties_and_edges" 2>&1 | grep -v "^$" | cut -c1-260 | tail -90
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
for F in 1 0; do
  DINER_TC_FUSED=$F timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_fused$F.json 2> gpurun_out/r2f_bench_fused$F.err
  python -c "
import json;d=json.load(open('gpurun_out/r2f_bench_fused$F.json'));print('fused',$F,d['value'],d['e2e']['value'],d['roofline']['frac'],d['roofline']['stage_ms_per_step'],d['clocks'],d['gpu_launches'],d['parity'])"
done
