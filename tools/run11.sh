cd $GRAFT_REPO_ROOT
export PYTHONFAULTHANDLER=1
timeout 900 python -m pytest tests/test_gpu_dense_keys.py tests/test_gpu_large.py -m gpu -q --timeout 240 > gpurun_out/r02_tests8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests8.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02_tests8.log | tail -20
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 1 --configs none"
PW_DEBUG=1 $B --steps 3 2>&1 | grep "bucket tier" | head -1
run() { name=$1; shift; env "$@" $B > gpurun_out/r02_b7_$name.json 2> gpurun_out/r02_b7_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_b7_$name.json").read().strip().splitlines()[-1])
    print("$name", "step", round(d["ms_per_step"],4), "kernel", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["frac"],4), "spilled", d.get("spilled_rows"), d.get("check"))
except Exception as e:
    print("$name", "ERR", e, open("gpurun_out/r02_b7_$name.err").read()[-800:])
PY
}
run s3 X=1
run s2 PW_BUCKET_STAGES=2
run s3v1 PW_BUCKET_VAR=1
run s3v2 PW_BUCKET_VAR=2
run s3v3 PW_BUCKET_VAR=3
run s0 PW_BUCKET_STAGES=0
run s0v3 PW_BUCKET_STAGES=0 PW_BUCKET_VAR=3
run s4j6 PW_BUCKET_STAGES=4 PW_BUCKET_J=6
