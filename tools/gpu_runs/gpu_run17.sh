cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --tb=short -k "tile_order or whole_image or sweep" 2>&1 | tail -4
for Wd in 512 0; do
  DINER_RAY_IMAGE_WIDTH=$Wd timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2o_bench_w$Wd.json 2> gpurun_out/r2o_bench_w$Wd.err
  python -c "
import json;d=json.load(open('gpurun_out/r2o_bench_w$Wd.json'));print('ray_image_width',$Wd,d['value'],d['e2e']['value'],d['roofline']['frac'],d['roofline']['stage_ms_per_step'],d['clocks'])"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_pair_kernel -s 1 -c 1 -o gpurun_out/r2o_fused_tiled python tools/profile_run.py parity 8192 2 > gpurun_out/r2o_ncu_full.log 2>&1; tail -2 gpurun_out/r2o_ncu_full.log
