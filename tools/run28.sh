cd $GRAFT_REPO_ROOT
export PYTHONFAULTHANDLER=1
timeout 200 python -m pytest tests/test_gpu_boundary.py -m gpu -q --timeout 100 2>&1 | tail -3
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --configs none"
timeout 200 $B > gpurun_out/plain4.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c2_bench_launches_ncu.csv $B > gpurun_out/ncu7.log 2>&1
tail -2 gpurun_out/ncu7.log
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02_c2_bench_launches_ncu.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
seq=[(r[ki][:60], float(r[vi].replace(",",""))) for r in rows[1:]]
print(len(seq))
for k,v in seq[-24:]: print(f"{v:10.1f}  {k}")
PY
