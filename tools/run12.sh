cd $GRAFT_REPO_ROOT
export PYTHONFAULTHANDLER=1
timeout 900 python -m pytest tests/test_gpu_dense_keys.py -m gpu -q --timeout 240 > gpurun_out/r02_tests9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests9.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02_tests9.log | tail -20
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 1 --configs none"
PW_DEBUG=1 $B --steps 3 2>&1 | grep "bucket tier" | head -1
run() { name=$1; shift; env "$@" $B > gpurun_out/r02_b8_$name.json 2> gpurun_out/r02_b8_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_b8_$name.json").read().strip().splitlines()[-1])
    print("$name", "step", round(d["ms_per_step"],4), "kernel", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["frac"],4), "spilled", d.get("spilled_rows"), d.get("check"))
except Exception as e:
    print("$name", "ERR", e, open("gpurun_out/r02_b8_$name.err").read()[-800:])
PY
}
run nobucket PW_NO_BUCKET=1
run s2j9 X=1
run s2j9v2 PW_BUCKET_VAR=2
run s2j8 PW_BUCKET_J=8
run s0j9 PW_BUCKET_STAGES=0 PW_BUCKET_J=9
run s0j10 PW_BUCKET_STAGES=0
run c512 PW_BUCKET_CAND=1
run nobucket2 PW_NO_BUCKET=1
