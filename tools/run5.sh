cd $GRAFT_REPO_ROOT
timeout 300 ./tools/bucket_probe > gpurun_out/r02_bucket_probe2.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_bucket_probe2.txt
cat gpurun_out/r02_bucket_probe2.txt
