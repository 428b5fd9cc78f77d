# bench line after the traffic-summary change (default workload)
cd $GRAFT_REPO_ROOT
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2zz_err.txt | grep "^{" > gpurun_out/r2zz_bench2.json; python -c "
import json;d=json.load(open('gpurun_out/r2zz_bench2.json'));r=d['roofline'];print(d['value'],d['e2e']['value'],r['frac'],r['traffic'],r['traffic_captured'],r['traffic_source'],r['achieved_executed_mma'])" || tail -5 gpurun_out/r2zz_err.txt
