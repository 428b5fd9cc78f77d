# N = 2 with the final kernel: sharded image == single-GPU image, bench line
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
N=$1
timeout 300 $TR --nproc-per-node $N --master-port 2951$N tools/check_sharded_equal.py 2>&1 | tail -2 | cut -c1-300
timeout 600 $TR --nproc-per-node $N --master-port 2952$N bench.py --gpus $N --steps 5 --warmup 3 2>/dev/null | grep "^{" > gpurun_out/r2z_bench_${N}gpu.json
python -c "
import json;d=json.load(open('gpurun_out/r2z_bench_${N}gpu.json'));print('N',$N,d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['stage_ms_per_step'],d['clocks'])"
