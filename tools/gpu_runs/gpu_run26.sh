# option sweep on the final kernel (results are bit-identical for all of these; only the time changes)
cd $GRAFT_REPO_ROOT
run() { env "$@" timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | grep "^{" | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$*',d['value'],d['roofline']['frac'],d['clocks']['sm_mhz'])"; }
run A=0
run DINER_TC_TAIL_KB=2
run DINER_TC_TAIL_KB=4
run DINER_TC_EARLY_SPLIT=3
run DINER_TC_EARLY_SPLIT=5
run A=1
