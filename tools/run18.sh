cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -3
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err ) 2>&1 | tail -3
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_n2.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","e2e","gpu_launches","n_gpus","check")})
    print("roofline", d["roofline"])
    for k,v in d["configs"].items():
        print(k, {x: v.get(x) for x in ("rows","ms_per_step","rows_per_s","strategy","kernel_ms","frac_measured","n_groups","check","error","nccl_ms_per_step","exchanged_bytes_per_rank")})
except Exception as e:
    print("ERR", e)
PY
tail -8 gpurun_out/r02_bench_n2.err
timeout 600 python -m pytest tests/test_gpu_multigpu.py -m gpu -q --timeout 240 2>&1 | tail -3
