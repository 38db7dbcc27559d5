cd $GRAFT_REPO_ROOT
nproc; free -g | head -2
( time python bench.py > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_full.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches","clocks")})
print("roofline", d["roofline"]); print("cpu", d["cpu_baseline"]); print("check", d.get("check"))
for k,v in d["configs"].items():
    print(k, {x: v.get(x) for x in ("rows","ms_per_step","strategy","kernel","kernel_ms","frac_measured","whole_step_frac_measured","n_groups","check","error")})
PY
tail -5 gpurun_out/r02_bench_full.err
( time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err ) 2>&1 | tail -3
cat gpurun_out/r02_bench_ref.json | cut -c1-800
