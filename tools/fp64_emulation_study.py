#!/usr/bin/env python
"""Go / no-go numbers for an FP64-emulating DGEMM on the integer tensor cores - numpy only, nothing here runs on a GPU
and nothing of it is in the product.

Scheme (Ozaki-type, error-free slicing): every row i of A is scaled by 2^-ea[i] (ea[i] = exponent of the row's largest magnitude) and
cut into S signed slices of B bits,  a[i,l] = 2^ea[i] * sum_s 2^-(B - 1 + B s) * qa_s[i,l] + remainder,  |qa_s| <= 2^(B-1) (fits int8 for
B <= 7; round-to-nearest slicing of the running remainder); columns of B likewise.  The slice products  qa_s * qb_t  are exact integer GEMMs
(int8 x int8 -> int32 is exact for k <= 2^31 / 2^(2B-2), i.e. k <= 2^19 at B = 7; longer k is windowed), and
    C ~= sum over s + t < S of  2^(ea[i] + eb[j] - 2 (B - 1) - B (s + t)) * (Qa_s Qb_t)[i, j]          S (S + 1) / 2 integer GEMMs,
summed in FP64 from the smallest terms up.  Error: the dropped terms s + t >= S, about (S + 1) 2^-(B (S + 1)) per product relative to the
row / column maxima, plus FP64 rounding of the final sum.

The tests' bound is  max|C - C_ref| / (k max|A| max|B|) <= 1e-15  (tests/_util.py TOL for float64); the study prints that ratio for
uniform(-1, 1) operands (the benchmark's data), for rows / columns graded over ten decades (where per-row scaling matters), and for
small integers (the reference's own test data: must be exact), together with the number of integer GEMMs and the FP64-equivalent rate that
number implies at a given fraction of the INT8 tensor peak (B200: 4.5e15 dense int8 op/s nominal)."""
import argparse

import numpy as np

INT8_PEAK_OPS = 4.5e15  # nominal dense int8 op/s (2 ops per MAC), /opt/skills/guides: fp8 = int8 rate on sm_100a


def slice_rows(x, bits, count):
    """x (rows x k) -> (exponents[rows], [count] integer slices, p0) with x ~= 2^e * sum_s 2^-(p0 + bits s) q_s, |q_s| <= 2^(bits-1):
    round-to-nearest slicing of the running remainder; the first slice carries bits - 1 magnitude bits (p0 = bits - 1)."""
    amax = np.max(np.abs(x), axis=1)
    e = np.where(amax > 0, np.floor(np.log2(np.where(amax > 0, amax, 1.0))) + 1, 0).astype(np.int64)  # |x| < 2^e
    r = np.ldexp(x, -e[:, None])  # in (-1, 1), exact
    p0 = bits - 1
    out = []
    for s in range(count):
        p = p0 + bits * s
        q = np.rint(np.ldexp(r, p))
        r = r - np.ldexp(q, -p)  # exact: |r| <= 2^-(p+1) afterwards
        assert np.max(np.abs(q)) <= 2 ** (bits - 1), (s, np.max(np.abs(q)))
        out.append(q.astype(np.int64))
    return e, out, p0


def emulated_gemm(a, b, bits, count):
    ea, qa, s0 = slice_rows(a, bits, count)
    eb, qb, _ = slice_rows(b.T.copy(), bits, count)
    m, n = a.shape[0], b.shape[1]
    c = np.zeros((m, n), np.float64)
    gemms = 0
    for total in range(2 * count - 2, -1, -1):  # smallest terms first
        if total >= count:
            continue  # dropped: s + t >= S
        acc = np.zeros((m, n), np.int64)
        for s in range(total + 1):
            t = total - s
            acc += qa[s] @ qb[t].T  # exact integers (int64 here; int32 windows on the tensor core)
            gemms += 1
        shift = -(2 * s0 + bits * total)
        c += np.ldexp(acc.astype(np.float64), shift)  # |acc| < 2^53 for k <= 2^38: the conversion is exact
    return np.ldexp(c, (ea[:, None] + eb[None, :]).astype(np.int64)), gemms


def reference(a, b):
    """float128-accumulated reference where numpy has it (x86-64 long double: 64-bit mantissa), else compensated float64."""
    return (a.astype(np.longdouble) @ b.astype(np.longdouble))


def cases(m, n, k, seed):
    rng = np.random.default_rng(seed)
    yield "uniform(-1,1)", rng.random((m, k)) * 2 - 1, rng.random((k, n)) * 2 - 1
    grade_a = 10.0 ** rng.uniform(-5, 5, (m, 1))
    grade_b = 10.0 ** rng.uniform(-5, 5, (1, n))
    yield "rows/cols graded 1e-5..1e5", (rng.random((m, k)) * 2 - 1) * grade_a, (rng.random((k, n)) * 2 - 1) * grade_b
    yield "normal, heavy-tailed entries", rng.standard_normal((m, k)) * 10.0 ** rng.uniform(-3, 3, (m, k)), rng.standard_normal((k, n))
    yield "integers 0..9 (reference test data)", rng.integers(0, 10, (m, k)).astype(np.float64), rng.integers(0, 10, (k, n)).astype(np.float64)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=96)
    ap.add_argument("--n", type=int, default=80)
    ap.add_argument("--k", type=int, nargs="*", default=[1000, 10000])
    ap.add_argument("--bits", type=int, default=7)
    ap.add_argument("--slices", type=int, nargs="*", default=[5, 6, 7, 8, 9])
    ap.add_argument("--efficiency", type=float, default=0.5, help="assumed fraction of the int8 tensor peak the slice GEMMs reach")
    args = ap.parse_args()
    print(f"slice width {args.bits} bits; bound of the FP64 parity tests: 1.0e-15; DMMA DGEMM measured: 35.4 TF; PCIe roofline at 10000^3: 46 TF")
    for k in args.k:
        for name, a, b in cases(args.m, args.n, k, 3):
            ref = reference(a, b)
            scale = k * np.max(np.abs(a)) * np.max(np.abs(b))
            native = float(np.max(np.abs((a @ b).astype(np.longdouble) - ref)) / scale)
            row = [f"k={k:6d} {name:36s} native FP64 {native:.1e} |"]
            for S in args.slices:
                c, gemms = emulated_gemm(a, b, args.bits, S)
                err = float(np.max(np.abs(c.astype(np.longdouble) - ref)) / scale)
                row.append(f"S={S}: {err:.1e}")
            print(" ".join(row))
    print()
    for S in args.slices:
        g = S * (S + 1) // 2
        eff = INT8_PEAK_OPS * args.efficiency / g * 1e-12
        print(f"S={S}: {g:2d} int8 GEMMs per DGEMM -> {eff:6.1f} TF FP64-equivalent at {args.efficiency:.0%} of the int8 peak"
              f" (+ slicing / recombination passes: O(S (mk + kn) + S^2 mn / window) bytes through HBM)")
