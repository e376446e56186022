#!/usr/bin/env python
"""The reference's DEFAULT call mode (pin_host_buffers = true, tiled_mm.hpp:79) on pageable host memory: how long one call takes when the
three matrices are page-locked for the duration of the call, for this library (TMM_PIN_THREADS pieces registered concurrently; 1 = one
cudaHostRegister per matrix like the reference) and for the unmodified reference.  GPU box; development tool.
    TMM_PIN_THREADS=1 python tools/pin_study.py ; TMM_PIN_THREADS=8 python tools/pin_study.py"""
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import tiled_mm_b200 as tmm  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
rng = np.random.default_rng(0)
a, b, c = (rng.random(n * n) - 0.5 for _ in range(3))           # pageable numpy memory
with tmm.make_context(np.float64) as ctx:
    best = 1e9
    for r in range(4):
        t0 = time.perf_counter()
        tmm.gemm(ctx, "N", "N", n, n, n, 1.0, a, n, b, n, 0.0, c, n, pin_host_buffers=True, copy_c_back=True)
        dt = time.perf_counter() - t0
        best = min(best, dt) if r else best
        print(f"  ours  TMM_PIN_THREADS={os.environ.get('TMM_PIN_THREADS', '1')}: run {r}: {dt * 1e3:.1f} ms")
    print(f"ours, pin_host_buffers=true on pageable memory, n={n}: best {best * 1e3:.1f} ms = {2.0 * n ** 3 / best * 1e-12:.2f} TFLOP/s")
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
