# compute-sanitizer racecheck (shared-memory hazards) over the MLP kernel on the smoke-sized render
cd $GRAFT_REPO_ROOT
timeout 500 compute-sanitizer --tool racecheck --racecheck-report analysis --kernel-name kernel_substring=mlp_pair --print-limit 30 \
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_racecheck.log 2>&1
echo "rc=$?"; grep -v "^$" gpurun_out/r2z_racecheck.log | tail -40 | cut -c1-400
