"""NCCL all-gather timing and transport on this box (run under torchrun)."""
import os, time, torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
for rows in (5_000, 23_000):
    x = torch.full((rows, 100), rank, dtype=torch.int32, device="cuda")
    out = torch.empty((world * rows, 100), dtype=torch.int32, device="cuda")
    for _ in range(5):
        dist.all_gather_into_tensor(out, x)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        dist.all_gather_into_tensor(out, x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    if rank == 0:
        print(f"all_gather_into_tensor world={world} {rows} rows x 400 B per rank: {ms:.3f} ms  ({world * rows * 400 / ms / 1e6:.1f} GB/s out per rank)", flush=True)
idx = torch.randperm(40_000, device="cuda")
src = torch.empty((40_000, 100), dtype=torch.int32, device="cuda"); dst = torch.empty_like(src)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    dst.index_copy_(0, idx, src.index_select(0, idx))
e1.record(); torch.cuda.synchronize()
if rank == 0:
    print(f"index_select + index_copy of 40k x 400 B: {e0.elapsed_time(e1) / 20:.3f} ms", flush=True)
    print("can_device_access_peer(0,1):", torch.cuda.can_device_access_peer(0, 1) if world > 1 else None, flush=True)
dist.destroy_process_group()
