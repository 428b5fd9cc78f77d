# compute-sanitizer synccheck over the MLP kernel on the smoke-sized render
cd $GRAFT_REPO_ROOT
timeout 400 compute-sanitizer --tool synccheck --kernel-name kernel_substring=mlp_pair --print-limit 10 \
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_synccheck.log 2>&1
echo "rc=$?"; grep -v "^$" gpurun_out/r2z_synccheck.log | tail -25 | cut -c1-300
