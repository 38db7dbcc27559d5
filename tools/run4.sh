set -x
cd $GRAFT_REPO_ROOT
timeout 300 ./tools/bucket_probe > gpurun_out/r02_bucket_probe1.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_bucket_probe1.txt
cat gpurun_out/r02_bucket_probe1.txt
export PYTHONFAULTHANDLER=1
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 240 > gpurun_out/r02_tests3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests3.log
tail -30 gpurun_out/r02_tests3.log
