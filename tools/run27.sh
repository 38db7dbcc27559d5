cd $GRAFT_REPO_ROOT
nvidia-smi topo -m 2>&1 | head -8
NCCL_DEBUG=INFO timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/nccl_probe.py > gpurun_out/r02_nccl_probe.log 2>&1
grep -E "all_to_all|all_gather|via|NVLS|P2P|SHM|NET/" gpurun_out/r02_nccl_probe.log | grep -v "^$" | sed 's/^.*NCCL INFO //' | sort | uniq -c | sort -rn | head -20
