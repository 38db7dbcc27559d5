cd $GRAFT_REPO_ROOT
export PYTHONFAULTHANDLER=1
timeout 900 python -m pytest tests/test_gpu_sorted_keys.py tests/test_gpu_filter_groups.py -m gpu -q --timeout 240 > gpurun_out/r02_tests12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests12.log
grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/r02_tests12.log | tail -30
python - <<'PY'
import time, numpy as np, torch, sys, pyarrow as pa
sys.path.insert(0, ".")
import polaroid_b200 as pw
from polaroid_b200 import engine
rng = np.random.default_rng(1)
n = 100_000_000
for groups in (1000, 10_000_000):
    keys = np.sort(rng.integers(0, groups, n))
    t = pa.table({"key": pa.array(keys), "value": pa.array(rng.random(n))})
    f = engine.DeviceFrame(t)
    for name, lf in (("flagged", pw.LazyFrame(t).set_sorted("key")), ("unflagged", pw.LazyFrame(t))):
        plan = lf.group_by("key").agg(pw.col("value").sum().alias("s"), pw.col("value").mean().alias("m"), pw.col("value").min().alias("lo"), pw.col("value").max().alias("hi")).plan
        for i in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            out = f.group_by(plan)
            torch.cuda.synchronize(); t1 = time.perf_counter()
        tm = engine.last_timings()
        print(f"sorted keys, {groups} groups, {name}: wall {1e3*(t1-t0):.2f} ms strategy {tm['strategy']} scan_kernel {tm['scan_kernel_ms']:.3f} ms scan {tm['scan_ms']:.3f} ms -> {n*16/tm['scan_kernel_ms']/1e6:.0f} GB/s, groups {out.num_rows}", flush=True)
    f.free()
PY
