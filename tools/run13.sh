cd $GRAFT_REPO_ROOT
export PYTHONFAULTHANDLER=1
timeout 900 python -m pytest tests/test_gpu_windowed_bucket.py tests/test_gpu_dynamic.py tests/test_gpu_dense_keys.py -m gpu -q --timeout 240 > gpurun_out/r02_tests10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests10.log
grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/r02_tests10.log | tail -30
run() { name=$1; shift; env "$@" python bench.py --no-cpu-baseline --workload c4 --scale 0.2 --config-steps 3 > gpurun_out/r02_b9_$name.json 2> gpurun_out/r02_b9_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_b9_$name.json").read().strip().splitlines()[-1])["detail"]
    print("$name", {k: d.get(k) for k in ("rows","ms_per_step","strategy","kernel","kernel_ms","frac_measured","n_groups","check","error")}, d.get("phases_ms"))
except Exception as e:
    print("$name", "ERR", e, open("gpurun_out/r02_b9_$name.err").read()[-1500:], open("gpurun_out/r02_b9_$name.json").read()[-600:])
PY
}
run c4_win X=1
run c4_hash PW_NO_WBUCKET=1
run c4_cand1 PW_BUCKET_CAND=1
run c4_cand2 PW_BUCKET_CAND=2
run c4_s0 PW_BUCKET_STAGES=0
