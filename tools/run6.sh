cd $GRAFT_REPO_ROOT
timeout 300 ./tools/bucket_probe > gpurun_out/r02_bucket_probe3.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_bucket_probe3.txt
grep "^bucket\|^stream\|rc=" gpurun_out/r02_bucket_probe3.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bucket_kernel -s 5 -c 1 -o gpurun_out/r02_bucket_p3 ./tools/bucket_probe > gpurun_out/ncu_bp.log 2>&1
tail -3 gpurun_out/ncu_bp.log
