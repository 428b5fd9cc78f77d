# ncu --set full of the fused kernel at the stress shape (1024x1024, 8 views, 256 samples/ray): 16 image rows = 16 384 rays
cd $GRAFT_REPO_ROOT
timeout 420 ncu --set full --clock-control none --import-source on -k regex:mlp_pair_kernel -s 1 -c 1 -o gpurun_out/r2z_fused_stress python tools/profile_run.py parity 16384 2 stress1024 > gpurun_out/r2z_ncu_stress.log 2>&1; tail -3 gpurun_out/r2z_ncu_stress.log | cut -c1-300
