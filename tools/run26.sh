cd $GRAFT_REPO_ROOT
B="timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 --configs c2_hash --config-steps 20"
run() { name=$1; shift; env "$@" $B > gpurun_out/r02_b13_$name.json 2> gpurun_out/r02_b13_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_b13_$name.json").read().strip().splitlines()[-1])
    v=d["configs"]["c2_hash"]
    print("$name", {x: v.get(x) for x in ("ms_per_step","strategy","kernel_ms","frac_measured","error")})
except Exception as e: print("$name ERR", e)
PY
}
run mul4 X=1
run mul8 PW_BUCKET_IDXMUL=8
run old PW_NO_BUCKET_INDEX=1
run mul8j9 PW_BUCKET_IDXMUL=8 PW_BUCKET_STAGES=0 PW_BUCKET_J=9
