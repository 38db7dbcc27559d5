"""dev probe: all_to_all_single / all_gather bandwidth and latency between the ranks of one box (torchrun)."""
import os, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl")
dev = torch.device("cuda", lr)
for mb in (1, 64, 800):
    n = mb * 1024 * 1024 // 8 // world * world
    send = torch.ones(n, dtype=torch.int64, device=dev); recv = torch.empty_like(send)
    for _ in range(3): dist.all_to_all_single(recv, send)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): dist.all_to_all_single(recv, send)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    if rank == 0: print(f"all_to_all_single {mb} MB per rank: {ms:.3f} ms -> {n * 8 * (world - 1) / world / ms / 1e6:.1f} GB/s sent per rank", flush=True)
small = torch.ones(8192 * 8, dtype=torch.int64, device=dev); out = torch.empty(world * small.numel(), dtype=torch.int64, device=dev)
for _ in range(5): dist.all_gather_into_tensor(out, small)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): dist.all_gather_into_tensor(out, small)
e1.record(); torch.cuda.synchronize()
if rank == 0: print(f"all_gather_into_tensor 512 KB per rank: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us", flush=True)
# uneven splits as the partial exchange uses them
n = 100_000_000 // world * world
send = torch.ones(n, dtype=torch.int64, device=dev)
splits = [n // world] * world
recv = torch.empty(n, dtype=torch.int64, device=dev)
for _ in range(2): dist.all_to_all_single(recv, send, output_split_sizes=splits, input_split_sizes=splits)
torch.cuda.synchronize(); dist.barrier()
e0.record()
for _ in range(3): dist.all_to_all_single(recv, send, output_split_sizes=splits, input_split_sizes=splits)
e1.record(); torch.cuda.synchronize()
if rank == 0: print(f"all_to_all_single with split lists, 800 MB per rank: {e0.elapsed_time(e1) / 3:.3f} ms", flush=True)
dist.destroy_process_group()
