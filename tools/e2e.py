#!/usr/bin/env python
"""Host-to-host timing of one configuration (GPU box; development tool).  Under torchrun it runs the grid path.
   python tools/e2e.py [--m M --n N --k K --tt NN --dtype s|d|c|z --beta 0 --reps 6 --trace]"""
import argparse, os, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=10000); ap.add_argument("--n", type=int, default=10000); ap.add_argument("--k", type=int, default=10000)
ap.add_argument("--tt", default="NN"); ap.add_argument("--dtype", default="d"); ap.add_argument("--beta", type=float, default=0.0)
ap.add_argument("--reps", type=int, default=6); ap.add_argument("--trace", action="store_true"); ap.add_argument("--budget", type=float, default=0.0)
ap.add_argument("--copy-c-back", type=int, default=1); ap.add_argument("--devices", type=int, default=0)
ap.add_argument("--fill", default="rand", help="rand | const (timing only: skips the slow host RNG)")
ap.add_argument("--trace-dir", default="", help="every rank traces (TMM_TRACE=1) into <dir>/trace_rank<r>.txt")
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
if args.trace and rank == 0:
    os.environ["TMM_TRACE"] = "1"
if args.trace_dir:
    os.environ["TMM_TRACE"] = "1"
    os.makedirs(args.trace_dir, exist_ok=True)
    _fd = os.open(os.path.join(args.trace_dir, f"trace_rank{rank}.txt"), os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
    os.dup2(_fd, 2)
import tiled_mm_b200 as tmm
dist = None
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
dt = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}[args.dtype]
m, n, k = args.m, args.n, args.k
ta, tb = args.tt
ar, ac = (m, k) if ta == "N" else (k, m)
br, bc = (k, n) if tb == "N" else (n, k)
a = tmm.malloc_pinned(dt, ar * ac); b = tmm.malloc_pinned(dt, br * bc); c = tmm.malloc_pinned(dt, m * n)
rng = np.random.default_rng(rank)
for arr in (a, b):
    if args.fill == "const":
        np.asarray(arr).view(np.float32 if args.dtype in "sc" else np.float64)[:] = 0.5
        continue
    v = arr.view(np.float32 if args.dtype in "sc" else np.float64)
    for off in range(0, v.size, 1 << 24):
        v[off:off + (1 << 24)] = rng.random(min(1 << 24, v.size - off)) - 0.5
ctx = tmm.make_context(dt, 2, 5000, 5000, 5000)
if args.budget:
    ctx.set_device_budget(int(args.budget * (1 << 30)))
if args.devices > 1:
    ctx.set_devices(args.devices)
grid = None
if world > 1:
    from tiled_mm_b200 import multi_gpu
    grid = multi_gpu.GridGemm(ctx, dist)
flops = (2.0 if args.dtype in "sd" else 8.0) * m * n * k * world
best = 1e30
for r in range(args.reps):
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    tmm.gemm(ctx, ta, tb, m, n, k, 1.0, a, ar, b, br, args.beta, c, m, pin_host_buffers=False, copy_c_back=bool(args.copy_c_back))
    if dist is not None:
        dist.barrier()
    dtm = (time.perf_counter() - t0) * 1e3
    best = min(best, dtm)
    st = ctx.last_stats()
    if rank == 0:
        print(f"  run {r}: {dtm:8.2f} ms  regime {st.regime} blocks {st.c_blocks} chunks {st.k_chunks} launches {st.kernel_launches} h2d {st.h2d_bytes/1e6:.0f} MB d2h {st.d2h_bytes/1e6:.0f} MB peer {st.peer_bytes/1e6:.0f} MB", flush=True)
if rank == 0:
    print(f"E2E {args.dtype}gemm {args.tt} {m}x{n}x{k} beta={args.beta} world={world} devices={max(1,args.devices)}: best {best:.2f} ms = {flops/best*1e-9:.2f} TFLOP/s", flush=True)
ctx.close()
if dist is not None:
    dist.destroy_process_group()
