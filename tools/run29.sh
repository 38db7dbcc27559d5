cd $GRAFT_REPO_ROOT
export PYTHONFAULTHANDLER=1
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r02_tests_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests_full.log
grep -E "^(FAILED|ERROR)|passed|failed|Error|rc=" gpurun_out/r02_tests_full.log | tail -20
