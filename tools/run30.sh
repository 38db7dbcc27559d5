cd $GRAFT_REPO_ROOT
export PYTHONFAULTHANDLER=1
timeout 600 python -m pytest tests/test_gpu_dense_keys.py tests/test_gpu_windowed_bucket.py tests/test_gpu_indexed_bucket.py tests/test_gpu_dynamic.py -m gpu -q --timeout 200 > gpurun_out/r02_tests16.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests16.log
grep -E "^(FAILED|ERROR)|passed|failed|Error|rc=" gpurun_out/r02_tests16.log | tail -20
run() { name=$1; shift; env "$@" timeout 200 python bench.py --no-cpu-baseline --workload c4 --scale 0.2 --config-steps 5 > gpurun_out/r02_b14_$name.json 2> gpurun_out/r02_b14_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_b14_$name.json").read().strip().splitlines()[-1])["detail"]
    print("$name", {k: d.get(k) for k in ("rows","ms_per_step","strategy","kernel_ms","frac_measured","check","error")})
except Exception as e:
    print("$name", "ERR", e, open("gpurun_out/r02_b14_$name.err").read()[-800:])
PY
}
PW_DEBUG=1 timeout 100 python bench.py --no-cpu-baseline --workload c4 --scale 0.02 --config-steps 2 2>&1 | grep "bucket tier" | head -1
run rowpos X=1
run meta PW_NO_ROWPOS=1
run rowpos_c2 PW_BUCKET_CAND=2
