cd $GRAFT_REPO_ROOT
for st in 2 6; do
python bench.py --no-cpu-baseline --workload c3 --config-steps $st > gpurun_out/r02_c3_$st.json 2> gpurun_out/r02_c3_$st.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_c3_$st.json").read().strip().splitlines()[-1])["detail"]
print("c3 steps=$st", d.get("ms_per_step"), d.get("phases_ms"), d.get("error"))
PY
done
python - <<'PY'
import time, numpy as np, torch, sys
sys.path.insert(0, ".")
import bench
from polaroid_b200 import engine
cols, tensors = bench.gen_c3_device(100_000_000, 10_000_000, seed=3)
f = engine.DeviceFrame.from_device(cols)
plan = bench.c3_plan()
for i in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = f.group_by(plan)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    tm = engine.last_timings()
    print(i, "wall ms", round((t1 - t0) * 1e3, 1), {k: round(tm[k], 2) for k in ("scan_ms", "finalize_ms", "d2h_ms", "total_device_ms", "host_ms") if k in tm}, out.num_rows)
PY
