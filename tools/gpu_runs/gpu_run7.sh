cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --tb=short -k "sweep or cam_sweep or properties_and_edges" 2>&1 | tail -6
for P in 1 2 4; do
  DINER_TC_POST_TILES=$P timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_pts$P.json 2> gpurun_out/r2g_bench_pts$P.err
  python -c "
import json;d=json.load(open('gpurun_out/r2g_bench_pts$P.json'));print('post_tiles',$P,d['value'],d['e2e']['value'],d['roofline']['frac'],d['roofline']['stage_ms_per_step'],d['clocks'],d['gpu_launches'])"
done
