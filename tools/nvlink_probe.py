"""NVLink / NCCL transport probe (GPU box, >= 2 GPUs).  torchrun: NCCL all-gather bandwidth + transport lines; plain: P2P and D2D copy bandwidth."""
import os, sys, time
import torch
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
if world > 1:
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    for mb in (8, 64, 256):
        n = mb << 20
        src = torch.empty(n, dtype=torch.uint8, device="cuda"); dst = torch.empty(n * world, dtype=torch.uint8, device="cuda")
        ms = timeit(lambda: dist.all_gather_into_tensor(dst, src))
        if rank == 0: print(f"NCCL all_gather {mb} MB/rank x{world}: {ms:.3f} ms -> recv {(world-1)*n/ms*1e-6:.1f} GB/s per rank", flush=True)
    dist.destroy_process_group()
else:
    nd = torch.cuda.device_count()
    n = 256 << 20
    a0 = torch.empty(n, dtype=torch.uint8, device="cuda:0"); b0 = torch.empty(n, dtype=torch.uint8, device="cuda:0")
    ms = timeit(lambda: b0.copy_(a0)); print(f"D2D same device 256 MB: {ms:.3f} ms -> {n/ms*1e-6:.1f} GB/s (read+write {2*n/ms*1e-6:.1f})")
    rows = n // 81920
    a2 = a0[: rows * 81920].view(rows, 81920)[:, :80000]; b2 = b0[: rows * 81920].view(rows, 81920)[:, :80000]
    ms = timeit(lambda: b2.copy_(a2)); print(f"D2D strided (torch kernel) 2D: {ms:.3f} ms -> {a2.numel()/ms*1e-6:.1f} GB/s")
    if nd > 1:
        print("can access peer 0->1:", torch.cuda.can_device_access_peer(0, 1))
        b1 = torch.empty(n, dtype=torch.uint8, device="cuda:1")
        ms = timeit(lambda: b1.copy_(a0)); print(f"P2P copy 0->1 256 MB: {ms:.3f} ms -> {n/ms*1e-6:.1f} GB/s")
