cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29511 tools/check_sharded_equal.py 2>&1 | grep "sharded render"
for B in 1 0; do
DINER_BALANCE=$B timeout 600 $TR --nproc-per-node 8 --master-port 2951$B bench.py --gpus 8 --steps 5 --warmup 3 2>/dev/null | grep "^{" > gpurun_out/r2s_bench_8gpu_balance$B.json
python -c "
import json;d=json.load(open('gpurun_out/r2s_bench_8gpu_balance$B.json'));print('balance',$B,d['value'],d['e2e']['value'],d['ms_per_step'],d['run'].get('shard_weights'),d['roofline']['stage_ms_per_step'])"
done
