#!/usr/bin/env python
"""A priori accuracy of the two operand splits of the FP32-accurate tcgen05 SGEMM (csrc/gemm_f32_tc.cu), emulated in numpy:
   rn    : hi = round-to-nearest TF32 of x, lo = round-to-nearest TF32 of (x - hi)            (the default, validated on hardware)
   trunc : hi = x with the low 13 mantissa bits dropped (what the tensor core reads from raw FP32 bits), lo = RN_TF32(x - hi)
           (TMM_TC_SPLIT=trunc: the tile is not rewritten, one third less shared-memory traffic in the split stage)
Products hi*hi + hi*lo + lo*hi are exact in the tensor core; the study isolates the REPRESENTATION error (accumulation error is the same
for both and is handled by the windowed FP32 promotion).  Prints max / rms error of a k-term dot product relative to k * max|a| * max|b|."""
import numpy as np


def rn_tf32(x):
    u = x.astype(np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def trunc_tf32(x):
    return (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def split(x, mode):
    hi = rn_tf32(x) if mode == "rn" else trunc_tf32(x)
    lo = rn_tf32((x.astype(np.float32) - hi).astype(np.float32))
    return hi.astype(np.float64), lo.astype(np.float64)


def study(k, rows=256, seed=0, dist="uniform"):
    rng = np.random.default_rng(seed)
    gen = (lambda s: (rng.random(s) * 2 - 1)) if dist == "uniform" else (lambda s: rng.standard_normal(s))
    a = gen((rows, k)).astype(np.float32)
    b = gen((k, rows)).astype(np.float32)
    exact = a.astype(np.float64) @ b.astype(np.float64)
    out = {}
    for mode in ("rn", "trunc"):
        ah, al = split(a, mode)
        bh, bl = split(b, mode)
        approx = ah @ bh + ah @ bl + al @ bh
        err = np.abs(approx - exact)
        out[mode] = (err.max() / k, np.sqrt((err ** 2).mean()) / k)
    fp32_chain = np.abs((a.astype(np.float32) @ b.astype(np.float32)).astype(np.float64) - exact)   # numpy's own FP32 GEMM for scale
    out["fp32 gemm (numpy)"] = (fp32_chain.max() / k, np.sqrt((fp32_chain ** 2).mean()) / k)
    return out


if __name__ == "__main__":
    for k in (256, 4096, 32768):
        r = study(k)
        print(f"k = {k:6d}: " + " | ".join(f"{m}: max/k {v[0]:.2e} rms/k {v[1]:.2e}" for m, v in r.items()))
