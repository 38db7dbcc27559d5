set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests1.log
tail -5 gpurun_out/r02_tests1.log
./tools/smem_probe > gpurun_out/r02_smem_probe.txt 2>&1; cat gpurun_out/r02_smem_probe.txt
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/r02_b1_c2.json 2> gpurun_out/r02_b1_c2.err; cat gpurun_out/r02_b1_c2.json
PW_NO_DENSE=1 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/r02_b1_c2_nodense.json 2>&1
PW_NO_DEFERRED=1 PW_NO_PILOT_CACHE=1 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/r02_b1_c2_olddrive.json 2>&1
python bench.py --workload c1 --steps 30 --warmup 3 --e2e-steps 1 > gpurun_out/r02_b1_c1.json 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pw_scan_jit -s 3 -c 1 -o gpurun_out/r02_c2_v13 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu1.log 2>&1
tail -3 gpurun_out/ncu1.log
