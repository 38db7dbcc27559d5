cd $GRAFT_REPO_ROOT
export PYTHONFAULTHANDLER=1
timeout 900 python -m pytest tests/test_gpu_f1_keys.py tests/test_gpu_golden.py -m gpu -q --timeout 240 > gpurun_out/r02_tests13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests13.log
grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/r02_tests13.log | tail -30
