cd $GRAFT_REPO_ROOT
timeout 600 python tools/debug_tail.py 2>&1 | tail -40
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -s 2>&1 | grep -E "^(NV=|[a-z0-9_]+(/[a-z0-9]+)?:|full-size|whole-image|gen_rays|depth2normal|fast mode|training|[0-9]+ (passed|failed)|FAILED)" | cut -c1-330 > gpurun_out/r2b_pytest_prints.log
cat gpurun_out/r2b_pytest_prints.log
for T in 0 3; do
  DINER_TC_DBG_SKIP=512 DINER_TC_TAIL_KB=$T timeout 300 python tools/profile_run.py parity 8192 2 2>&1 | grep -E "ts\]|rep" | tee gpurun_out/r2b_ts_tail$T.log
done
