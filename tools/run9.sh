cd $GRAFT_REPO_ROOT
export PYTHONFAULTHANDLER=1
timeout 600 python -m pytest tests/test_gpu_dense_keys.py -m gpu -x -q --timeout 240 > gpurun_out/r02_tests6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests6.log
tail -3 gpurun_out/r02_tests6.log
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 1 --configs none"
$B > gpurun_out/r02_b5_bucket.json 2> gpurun_out/r02_b5_bucket.err
for f in bucket; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_b5_$f.json").read().strip().splitlines()[-1])
    print("$f", "step", round(d["ms_per_step"],4), "kernel", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["frac"],4), d["phases_ms"], d.get("check"))
except Exception as e:
    print("$f", "ERR", e, open("gpurun_out/r02_b5_$f.json").read()[-800:])
PY
done
ncu --set full --clock-control none --import-source on -k regex:pw_bucket_jit -s 3 -c 1 -o gpurun_out/r02_c2_bucket_v2 $B --steps 2 > gpurun_out/ncu3.log 2>&1
tail -3 gpurun_out/ncu3.log
