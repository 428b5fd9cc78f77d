# compute-sanitizer memcheck over the library's kernels on the smoke-sized render (parity + fast modes)
cd $GRAFT_REPO_ROOT
timeout 500 compute-sanitizer --tool memcheck --kernel-name kernel_substring=mlp_pair --kernel-name kernel_substring=sampler_kernel --kernel-name kernel_substring=composite_kernel --kernel-name kernel_substring=pack_ --print-limit 20 \
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_memcheck.log 2>&1
echo "rc=$?"; grep -v "^$" gpurun_out/r2z_memcheck.log | tail -12 | cut -c1-300
