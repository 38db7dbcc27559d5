set -x
cd $GRAFT_REPO_ROOT
export PYTHONFAULTHANDLER=1
cat > /tmp/conc.py <<'PY'
import faulthandler, sys
faulthandler.dump_traceback_later(90, exit=True)
import pytest
sys.exit(pytest.main(["tests/test_gpu_boundary.py", "-x", "-q", "-k", "concurrent"]))
PY
timeout 150 python /tmp/conc.py > gpurun_out/r02_conc.log 2>&1; echo "rc=$?" >> gpurun_out/r02_conc.log; tail -60 gpurun_out/r02_conc.log
