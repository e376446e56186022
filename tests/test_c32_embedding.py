"""CPU check of the index algebra behind the tcgen05 CGEMM (csrc/gemm_c32_tc.cu): the elementwise passes are restated here with
the same formulas (which float lands where) and the real product (2m x n) = (2m x 2k)(2k x n) over the (re, im) float views must
reproduce the complex GEMM of the oracle for every op pair, padded leading dimensions, complex alpha and real / complex beta.
The CUDA passes themselves have not run on hardware yet (the path is opt-in, TMM_C32_MATH=tc); this pins the mathematics."""
import itertools

import numpy as np
import pytest


def embed_a(ta, a, lda, m, k, alpha):
    """-> (stored float matrix as 2-D [col][row] array, op for the real GEMM)"""
    a = a.astype(np.complex64)
    if ta == "N":  # embed_a_n: column 2l = (re, im) of w(:, l); column 2l+1 = (-im, re)
        pitch = -(-2 * m // 32) * 32
        out = np.zeros((2 * k, pitch), np.float32)
        for l in range(k):
            w = (alpha * a[l * lda: l * lda + m]).astype(np.complex64)
            out[2 * l, 0:2 * m:2], out[2 * l, 1:2 * m:2] = w.real, w.imag
            out[2 * l + 1, 0:2 * m:2], out[2 * l + 1, 1:2 * m:2] = -w.imag, w.real
        return out, "N", pitch
    pitch = -(-2 * k // 32) * 32  # embed_a_t: stored k x m; out = A'^T (2k x 2m), k-contiguous
    out = np.zeros((2 * m, pitch), np.float32)
    for i in range(m):
        v = a[i * lda: i * lda + k]
        if ta == "C":
            v = np.conj(v)
        w = (alpha * v).astype(np.complex64)
        out[2 * i, 0:2 * k:2], out[2 * i, 1:2 * k:2] = w.real, -w.imag
        out[2 * i + 1, 0:2 * k:2], out[2 * i + 1, 1:2 * k:2] = w.imag, w.real
    return out, "T", pitch


def embed_b(tb, b, ldb, k, n):
    b = b.astype(np.complex64)
    if tb == "N":  # zero copy: the stored matrix read as floats, pitch 2 * ldb
        flat = b.view(np.float32)
        out = np.stack([flat[j * 2 * ldb: j * 2 * ldb + 2 * k] for j in range(n)])
        return out, "N"
    pitch = -(-n // 32) * 32  # split_b_t: stored n x k; out = B'^T (n x 2k), n-contiguous
    out = np.zeros((2 * k, pitch), np.float32)
    for l in range(k):
        v = b[l * ldb: l * ldb + n]
        out[2 * l, :n] = v.real
        out[2 * l + 1, :n] = -v.imag if tb == "C" else v.imag
    return out, "T"


def real_gemm(opa, sa, rows_a, opb, sb, m2, n, k2):
    A = sa[:, :rows_a].T if opa == "N" else sa[:, :rows_a]          # [col][row] storage -> matrix
    A = A[:m2, :k2] if opa == "N" else sa[:m2, :k2]
    B = sb[:, :k2].T if opb == "N" else sb[:k2, :n]
    B = B[:k2, :n]
    return A.astype(np.float64) @ B.astype(np.float64)


@pytest.mark.parametrize("ta,tb", list(itertools.product("NTC", "NTC")))
@pytest.mark.parametrize("beta", [0.0, 2.0, 1 - 2j])
def test_real_embedding_reproduces_complex_gemm(oracle, ta, tb, beta):
    rng = np.random.default_rng(hash((ta, tb)) % 1000)
    m, n, k = 7, 5, 6
    ar, ac = (m, k) if ta == "N" else (k, m)
    br, bc = (k, n) if tb == "N" else (n, k)
    lda, ldb, ldc = ar + 3, br + 2, m + 1   # ldb even or odd is irrelevant here: the zero-copy view is emulated for any ld
    def gen(count):
        return (rng.integers(-9, 10, count) + 1j * rng.integers(-9, 10, count)).astype(np.complex64)
    a, b, c = gen(lda * ac), gen(ldb * bc), gen(ldc * n)
    alpha = np.complex64(2 - 1j)
    expect = oracle.gemm(ta, tb, m, n, k, alpha, a, lda, b, ldb, np.complex64(beta), c.copy(), ldc)
    sa, opa, _ = embed_a(ta, a, lda, m, k, alpha)
    sb, opb = embed_b(tb, b, ldb, k, n)
    prod = real_gemm(opa, sa, 2 * m, opb, sb, 2 * m, n, 2 * k)    # (2m x n) real
    cf = c.copy()
    beta_r = np.float32(np.real(beta))
    if np.imag(beta) != 0:                                         # complex beta: scale first, accumulate with 1
        for j in range(n):
            cf[j * ldc: j * ldc + m] *= np.complex64(beta)
        beta_r = np.float32(1.0)
    out = cf.view(np.float32).astype(np.float64).reshape(n, 2 * ldc)
    for j in range(n):
        old = out[j, :2 * m] * float(beta_r) if beta_r != 0 else 0.0
        out[j, :2 * m] = prod[:, j] + old
    got = out.astype(np.float32).reshape(-1).view(np.complex64)
    assert np.array_equal(got, expect), (ta, tb, beta)
