cd $GRAFT_REPO_ROOT
B="python bench.py --no-cpu-baseline --workload c4 --scale 0.05 --config-steps 2"
$B > gpurun_out/plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pw_bucket_jit -s 2 -c 1 -o gpurun_out/r02_c4_wbucket_v1 $B > gpurun_out/ncu4.log 2>&1
tail -3 gpurun_out/ncu4.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --configs none"
$B > gpurun_out/plain2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pw_bucket_jit -s 3 -c 1 -o gpurun_out/r02_c2_bucket_v3 $B > gpurun_out/ncu5.log 2>&1
tail -3 gpurun_out/ncu5.log
