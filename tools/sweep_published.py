#!/usr/bin/env python
"""The reference's published experiment (README.md:20-25, docs/performance.svg: dgemm, square n = 4000 ... 32000, column-major,
alpha = beta = 1, pinned host buffers) re-run on this box for BOTH arms: this library and the unmodified reference + cuBLAS
(oracle/_ref/libtiledmm_ref.so, tile 5000^3 / 2 streams = its defaults).  Prints one table row per size: time, TFLOP/s, % of
min(FP64 peak, AI x PCIe), speed-up.  GPU box; development tool.

    python tools/sweep_published.py [--sizes 4000,8000,...] [--reps 3] [--no-reference]"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import tiled_mm_b200 as tmm  # noqa: E402

try:
    FP64_PEAK = tmm.probe_fp64_peak() * 1e12
    PCIE = tmm.probe_host_links([0], nbytes=256 << 20)[0][0] * 1e9
except Exception:
    FP64_PEAK, PCIE = 36.9e12, 55.6e9

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="4000,8000,12000,16000,20000,24000,28000,32000")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--beta", type=float, default=1.0)
ap.add_argument("--no-reference", action="store_true")
ap.add_argument("--dtype", default="d", help="s | d | c | z (the reference instantiates all four, tiled_mm.cpp:626-668)")
ap.add_argument("--copy-c-back", type=int, default=1, help="0: the miniapp's second variant - C stays on the device (examples/multiply.cpp:196-229)")
args = ap.parse_args()
sizes = [int(s) for s in args.sizes.split(",")]
nmax = max(sizes)
DT = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}[args.dtype]
FLOP_PER_MNK = 8.0 if args.dtype in "cz" else 2.0
a = tmm.malloc_pinned(DT, nmax * nmax); b = tmm.malloc_pinned(DT, nmax * nmax); c = tmm.malloc_pinned(DT, nmax * nmax)
real = np.float32 if args.dtype in "sc" else np.float64
slab = (np.random.default_rng(0).random(1 << 24) - 0.5).astype(real)   # one random slab repeated (the host RNG would take minutes for 25 GB)
for arr in (a, b, c):
    v = np.asarray(arr).view(real)
    for off in range(0, v.size, 1 << 24):
        v[off:off + (1 << 24)] = slab[:min(1 << 24, v.size - off)]

ref = None
if not args.no_reference:
    try:
        import _util
        ref = _util.Reference(cpu=False)
    except Exception as e:  # the reference .so did not travel
        print(f"reference arm unavailable: {e}")

ours = tmm.make_context(DT, 2, 5000, 5000, 5000)
theirs = ref.context(DT, 2, 5000, 5000, 5000) if ref else None
print(f"{args.dtype}gemm n x n x n, alpha = 1, beta = {args.beta}, copy_c_back = {bool(args.copy_c_back)}; FP64 peak {FP64_PEAK * 1e-12:.1f} TF, PCIe {PCIE * 1e-9:.1f} GB/s (both probed live)")
print(f"{'n':>6} | {'ours ms':>9} {'TF':>6} {'% roof':>6} | {'reference ms':>12} {'TF':>6} | speed-up | PCIe bytes ours / reference")
for n in sizes:
    flops = FLOP_PER_MNK * n ** 3
    def best(fn):
        fn()  # warm-up: context buffers grow on the first call of a size
        t = 1e30
        for _ in range(args.reps):
            t0 = time.perf_counter(); fn(); t = min(t, time.perf_counter() - t0)
        return t
    t_ours = best(lambda: tmm.gemm(ours, "N", "N", n, n, n, 1.0, a, n, b, n, args.beta, c, n, pin_host_buffers=False, copy_c_back=bool(args.copy_c_back)))
    st = ours.last_stats()
    moved = st.h2d_bytes + st.d2h_bytes
    roof = min(FP64_PEAK if args.dtype in "dz" else 1e30, flops / (moved / PCIE))   # float types: the PCIe bound only (no FP32 tensor peak is claimed)
    row = f"{n:>6} | {t_ours * 1e3:9.2f} {flops / t_ours * 1e-12:6.2f} {100 * flops / t_ours / roof:6.1f} | "
    if theirs:
        t_ref = best(lambda: theirs.gemm("N", "N", n, n, n, 1.0, a, n, b, n, args.beta, c, n, pin=False, copy_c_back=bool(args.copy_c_back)))
        tile = tmm.optimal_tile_size(n, 5000)
        nt = -(-n // tile)
        ref_bytes = np.dtype(DT).itemsize * n * n * (2 * nt + (2 if args.beta else 1))   # n_tiles_n |A| + n_tiles_m |B| + [beta] |C| up, |C| down
        row += f"{t_ref * 1e3:12.2f} {flops / t_ref * 1e-12:6.2f} | {t_ref / t_ours:7.2f}x | {moved / 1e9:.2f} GB / {ref_bytes / 1e9:.2f} GB"
    print(row, flush=True)
