import sys, time, numpy as np, pyarrow as pa, torch
sys.path.insert(0, ".")
import polaroid_b200 as pw
from polaroid_b200 import engine
rng = np.random.default_rng(2)
n = 100_000_000
mode = sys.argv[1] if len(sys.argv) > 1 else "dense"
keys = rng.integers(0, 1000, n)
if mode == "sparse": keys = keys * 7919
t = pa.table({"key": pa.array(keys), "value": pa.array(rng.random(n))})
f = engine.DeviceFrame(t)
plan = pw.LazyFrame(t).group_by("key").agg(pw.col("value").sum().alias("s"), pw.col("value").mean().alias("m"), pw.col("value").min().alias("lo"), pw.col("value").max().alias("hi")).plan
flags = engine.FLAG_NO_DENSE_IDS if mode == "nodense" else 0
for i in range(4):
    out = f.group_by(plan, flags=flags)
tm = engine.last_timings()
print(mode, {k: tm[k] for k in ("strategy", "scan_kernel_ms", "spilled_rows", "n_groups", "retries")}, flush=True)
