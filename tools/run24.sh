cd $GRAFT_REPO_ROOT
timeout 120 python tools/idx_probe.py sparse || echo "TIMEOUT/FAIL sparse"
timeout 120 python tools/idx_probe.py nodense || echo "TIMEOUT/FAIL nodense"
PW_BUCKET_IDXMUL=8 timeout 120 python tools/idx_probe.py sparse || echo "TIMEOUT/FAIL mul8"
export PYTHONFAULTHANDLER=1
timeout 300 python -m pytest tests/test_gpu_indexed_bucket.py -m gpu -q --timeout 100 > gpurun_out/r02_tests15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests15.log
grep -E "^(FAILED|ERROR)|passed|failed|Error|rc=" gpurun_out/r02_tests15.log | tail -30
