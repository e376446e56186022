#!/usr/bin/env python
"""Back-of-the-envelope timeline of the RESIDENT schedule of one call on one GPU (csrc/tmm_context.cu run_resident), from the plan the
library itself produces (tmm_plan_describe) and four measured rates.  A design aid for shapes that have not been measured yet -
never a benchmark.  Calibration: dgemm 10000^3 beta=0 -> 58.4 ms predicted, 58.6 ms measured (profiles/r1_final_gpu_suite.txt).

Resources: one H2D engine, one D2H engine (full duplex), the SMs as one work-conserving server that serves the phase-1 chunk chain first
and back-fills with phase-2 column blocks (stream priorities).  A launch costs flops / P plus the read-modify-write of its C block in HBM.

    python tools/model_resident.py [--sizes 4000,8000,...] [--beta 1]"""
import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import tiled_mm_b200 as tmm  # noqa: E402

P, H2D, D2H, HBM, LAT = 35.7e12, 55.3e9, 52.1e9, 6.5e12, 8e-6


def predict(m, n, k, beta, plan=None, es=8, flop_per_fma=2.0):
    p = plan or tmm.plan_describe(np.float64, "N", "N", m, n, k, beta != 0, True, 150 << 30)
    assert p["regime"] == 0, "resident regime only"
    n1, chunks, blocks = p["n1"], p["chunks"], p["blocks"]
    t = 0.0
    p1 = min(4, max(1, n1 // 1024))               # phase-1 column stripes (run_resident)
    ready1 = []                                   # phase-1 chunk c may start when its A / B rows have landed
    for ci, kc in enumerate(chunks):
        if beta and ci < p1:                      # stripe ci's share of C travels right before k-chunk ci
            t += es * m * (n1 / p1) / H2D + LAT
        t += es * (m * kc + kc * n1) / H2D + LAT
        ready1.append(t)
    if beta and len(chunks) < p1:
        t += es * m * (n1 / p1) * (p1 - len(chunks)) / H2D
    ready2 = []
    for nb in blocks:
        t += es * (k * nb + (m * nb if beta else 0)) / H2D + LAT
        ready2.append(t)
    h2d_end = t
    work1 = [flop_per_fma * m * n1 * kc / P + es * m * n1 * (2 if (beta or i) else 1) / HBM for i, kc in enumerate(chunks)]
    work2 = [flop_per_fma * m * nb * k / P + es * m * nb * (2 if beta else 1) / HBM for nb in blocks]
    # SM server: fluid, phase 1 first
    dt, now, i1, left1, done2, left2 = 2e-6, 0.0, 0, work1[0], [None] * len(blocks), list(work2)
    phase1_end = None
    while i1 < len(work1) or any(d is None for d in done2):
        if i1 < len(work1) and now >= ready1[i1]:
            left1 -= dt
            if left1 <= 0:
                i1 += 1
                if i1 < len(work1):
                    left1 = work1[i1]
                else:
                    phase1_end = now + dt
        else:
            for j in range(len(blocks)):
                if done2[j] is None and now >= ready2[j]:
                    left2[j] -= dt
                    if left2[j] <= 0:
                        done2[j] = now + dt
                    break
        now += dt
    # D2H: phase-1 block when its chain ends, then the phase-2 blocks as they finish
    t = phase1_end + es * m * n1 / D2H + LAT
    for j, nb in enumerate(blocks):
        t = max(t, done2[j]) + es * m * nb / D2H + LAT
    return t, h2d_end, phase1_end, p


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="2000,4000,5000,6000,8000,10000,12000,16000,20000,24000,32000")
    ap.add_argument("--beta", type=float, default=1.0)
    a = ap.parse_args()
    print(f"{'n':>6} {'pred ms':>8} {'TF':>6} | H2D busy until, phase 1 ends (ms) | floor = max(compute, H2D, D2H) ms | n1 / chunks / blocks")
    for n in (int(s) for s in a.sizes.split(",")):
        t, h, p1, p = predict(n, n, n, a.beta)
        fl = 2.0 * n ** 3
        floor = max(fl / P, 8 * n * n * (2 + (1 if a.beta else 0)) / H2D, 8 * n * n / D2H)
        print(f"{n:>6} {t * 1e3:8.2f} {fl / t * 1e-12:6.2f} | {h * 1e3:7.2f} {p1 * 1e3:7.2f} | {floor * 1e3:7.2f} | {p['n1']} / {len(p['chunks'])} / {len(p['blocks'])}")
