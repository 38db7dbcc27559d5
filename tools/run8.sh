cd $GRAFT_REPO_ROOT
export PYTHONFAULTHANDLER=1
timeout 900 python -m pytest tests/test_gpu_dense_keys.py tests/test_gpu_large.py -m gpu -x -q --timeout 240 > gpurun_out/r02_tests5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests5.log
tail -5 gpurun_out/r02_tests5.log
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 1 --configs none"
$B > gpurun_out/r02_b4_bucket.json 2> gpurun_out/r02_b4_bucket.err
PW_NO_NATIVE_MINMAX=1 $B > gpurun_out/r02_b4_image.json 2>&1
for c in 1 3; do PW_BUCKET_CAND=$c $B > gpurun_out/r02_b4_cand$c.json 2>&1; done
for f in bucket image cand1 cand3; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_b4_$f.json").read().strip().splitlines()[-1])
    print("$f", "step", round(d["ms_per_step"],4), "kernel", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["frac"],4), d["phases_ms"], d.get("check"))
except Exception as e:
    print("$f", "ERR", e, open("gpurun_out/r02_b4_$f.json").read()[-800:])
PY
done
tail -5 gpurun_out/r02_b4_bucket.err
