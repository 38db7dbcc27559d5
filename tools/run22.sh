cd $GRAFT_REPO_ROOT
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 --configs c2_hash"
run() { name=$1; shift; env "$@" PW_DEBUG=1 $B > gpurun_out/r02_b12_$name.json 2> gpurun_out/r02_b12_$name.err; grep "bucket tier" gpurun_out/r02_b12_$name.err | sort | uniq -c | head -3; python - <<PY
import json
d=json.loads(open("gpurun_out/r02_b12_$name.json").read().strip().splitlines()[-1])
v=d["configs"]["c2_hash"]
print("$name", {x: v.get(x) for x in ("ms_per_step","strategy","kernel_ms","frac_measured","error")})
PY
}
run mul4 X=1
run mul2 PW_BUCKET_IDXMUL=2
run mul8 PW_BUCKET_IDXMUL=8
run mul4s0 PW_BUCKET_STAGES=0
