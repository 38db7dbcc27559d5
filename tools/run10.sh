cd $GRAFT_REPO_ROOT
export PYTHONFAULTHANDLER=1
timeout 900 python -m pytest tests/test_gpu_dense_keys.py tests/test_gpu_large.py tests/test_gpu_f2.py -m gpu -q --timeout 240 > gpurun_out/r02_tests7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests7.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02_tests7.log | tail -40
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 1 --configs none"
PW_DEBUG=1 $B --steps 3 2>&1 | grep "bucket tier" | head -1
$B > gpurun_out/r02_b6_s2.json 2> gpurun_out/r02_b6_s2.err
PW_BUCKET_STAGES=3 $B > gpurun_out/r02_b6_s3.json 2>&1
PW_BUCKET_STAGES=2 PW_BUCKET_J=8 $B > gpurun_out/r02_b6_s2j8.json 2>&1
PW_BUCKET_STAGES=0 $B > gpurun_out/r02_b6_s0.json 2>&1
PW_BUCKET_STAGES=0 PW_BUCKET_J=8 $B > gpurun_out/r02_b6_s0j8.json 2>&1
for f in s2 s3 s2j8 s0 s0j8; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_b6_$f.json").read().strip().splitlines()[-1])
    print("$f", "step", round(d["ms_per_step"],4), "kernel", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["frac"],4), "spilled", d.get("spilled_rows"), d.get("check"))
except Exception as e:
    print("$f", "ERR", e, open("gpurun_out/r02_b6_$f.json").read()[-800:])
PY
done
tail -5 gpurun_out/r02_b6_s2.err
