cd $GRAFT_REPO_ROOT
python tools/idx_probe.py sparse > gpurun_out/plain3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pw_bucket_jit -s 2 -c 1 -o gpurun_out/r02_c2_idx_v1 python tools/idx_probe.py sparse > gpurun_out/ncu6.log 2>&1
tail -2 gpurun_out/ncu6.log
