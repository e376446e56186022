#!/usr/bin/env python
"""The reference's DEFAULT call mode (pin_host_buffers = true, tiled_mm.hpp:79) on pageable host memory: how long one call takes when the
three matrices are page-locked for the duration of the call (tiled_mm.cpp:529-554, 606-623) - this library, this library with the
registration cache (TMM_PIN_CACHE=1: registered once per context, released with it), and the unmodified reference.  GPU box; development tool.
    python tools/pin_study.py [n]"""
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import tiled_mm_b200 as tmm  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
slab = np.random.default_rng(0).random(1 << 22) - 0.5
def pageable():
    x = np.empty(n * n)
    for off in range(0, x.size, slab.size):
        x[off:off + slab.size] = slab[:min(slab.size, x.size - off)]
    return x
a, b, c = pageable(), pageable(), pageable()           # pageable numpy memory
for label, cache in (("one cudaHostRegister per matrix and call (the reference's behaviour)", "0"), ("TMM_PIN_CACHE=1", "1")):
    os.environ["TMM_PIN_CACHE"] = cache
    with tmm.make_context(np.float64) as ctx:
        best = 1e9
        for r in range(4):
            t0 = time.perf_counter()
            tmm.gemm(ctx, "N", "N", n, n, n, 1.0, a, n, b, n, 0.0, c, n, pin_host_buffers=True, copy_c_back=True)
            dt = time.perf_counter() - t0
            best = min(best, dt) if r else best
            print(f"  ours, {label}: run {r}: {dt * 1e3:.1f} ms")
        print(f"ours, pin_host_buffers=true on pageable memory, n={n}, {label}: best {best * 1e3:.1f} ms = {2.0 * n ** 3 / best * 1e-12:.2f} TFLOP/s")
os.environ.pop("TMM_PIN_CACHE", None)
try:
    import _util
    ref = _util.Reference(cpu=False)
    rctx = ref.context(np.float64, 2, 5000, 5000, 5000)
    best = 1e9
    for r in range(3):
        t0 = time.perf_counter()
        rctx.gemm("N", "N", n, n, n, 1.0, a, n, b, n, 0.0, c, n, pin=True, copy_c_back=True)
        dt = time.perf_counter() - t0
        best = min(best, dt) if r else best
    print(f"reference, pin_host_buffers=true, n={n}: best {best * 1e3:.1f} ms = {2.0 * n ** 3 / best * 1e-12:.2f} TFLOP/s")
    rctx.close()
except Exception as e:  # noqa: BLE001
    print("reference arm unavailable:", e)
