# last check of the committed state: whole GPU suite + default bench line
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | cut -c1-300
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | grep "^{" > gpurun_out/r2zz_bench.json; python -c "
import json;d=json.load(open('gpurun_out/r2zz_bench.json'));print(d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['traffic'],d['clocks'],d['gpu_launches'])"
