cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout 600 python bench.py > gpurun_out/r2t_bench_default.json 2> gpurun_out/r2t_bench_default.err; python -c "
import json;d=json.load(open('gpurun_out/r2t_bench_default.json'));print(d['value'],d['e2e'],d['roofline']['frac'],d['roofline']['traffic'],d['roofline']['stage_ms_per_step'],d['clocks'],d['gpu_launches'],d['parity'],d['cpu_baseline']['value'],d['gpu_eager_baseline']['value'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-300
