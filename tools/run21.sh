cd $GRAFT_REPO_ROOT
export PYTHONFAULTHANDLER=1
timeout 900 python -m pytest tests/test_gpu_indexed_bucket.py tests/test_gpu_dense_keys.py tests/test_gpu_parity.py -m gpu -q --timeout 240 > gpurun_out/r02_tests14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests14.log
grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/r02_tests14.log | tail -30
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 1 --configs c2_hash,c1 > gpurun_out/r02_b11.json 2> gpurun_out/r02_b11.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_b11.json").read().strip().splitlines()[-1])
print("c2", d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])
for k,v in d["configs"].items():
    print(k, {x: v.get(x) for x in ("rows","ms_per_step","strategy","kernel","kernel_ms","frac_measured","check","error")})
PY
tail -3 gpurun_out/r02_b11.err
