set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests2.log
tail -15 gpurun_out/r02_tests2.log
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 1 --configs none"
$B > gpurun_out/r02_b2_default.json 2> gpurun_out/r02_b2_default.err
PW_NO_GUARD=1 $B > gpurun_out/r02_b2_noguard.json 2>&1
PW_NO_DEDUP=1 $B > gpurun_out/r02_b2_nodedup.json 2>&1
PW_NO_DEDUP=1 PW_NO_GUARD=1 $B > gpurun_out/r02_b2_claims_noguard.json 2>&1
for f in default noguard nodedup claims_noguard; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_b2_$f.json").read().strip().splitlines()[-1])
    print("$f", "step", round(d["ms_per_step"],4), "kernel", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["frac"],4), d["phases_ms"])
except Exception as e:
    print("$f", "ERR", e, open("gpurun_out/r02_b2_$f.json").read()[-800:])
PY
done
python bench.py --steps 10 --warmup 3 --e2e-steps 2 --configs all --scale 0.05 > gpurun_out/r02_b2_cfg_small.json 2> gpurun_out/r02_b2_cfg_small.err; tail -c 6000 gpurun_out/r02_b2_cfg_small.json; tail -5 gpurun_out/r02_b2_cfg_small.err
$B --steps 2 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pw_scan_jit -s 3 -c 1 -o gpurun_out/r02_c2_v14 $B --steps 2 > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
