"""Test helpers: the oracle (CPU restatement + fixture generator), the compiled reference, error metrics.
Test infrastructure only - nothing here is imported by the product."""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
CODES = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.complex64): 2, np.dtype(np.complex128): 3}


def _c(ch: str) -> ctypes.c_char:
    return ctypes.c_char(ch.encode()[:1])


def _vp(a) -> ctypes.c_void_p:
    return ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(0)


def ensure_oracle_built() -> None:
    if not (ORACLE_DIR / "liboracle.so").exists():
        subprocess.run(["make", "-C", str(ORACLE_DIR), "all"], check=True, capture_output=True)


class Oracle:
    def __init__(self):
        ensure_oracle_built()
        self.lib = ctypes.CDLL(str(ORACLE_DIR / "liboracle.so"))
        i64 = ctypes.c_int64
        self.lib.oracle_gemm_ex.argtypes = [ctypes.c_int, ctypes.c_char, ctypes.c_char, i64, i64, i64, ctypes.c_void_p, ctypes.c_void_p, i64,
                                            ctypes.c_void_p, i64, ctypes.c_void_p, ctypes.c_void_p, i64, ctypes.c_int]
        self.lib.oracle_fixture_fill.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]
        self.lib.oracle_fixture_reset.argtypes = [ctypes.c_uint]
        self.lib.oracle_fixture_lemire.argtypes = [ctypes.c_uint, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_size_t]
        self.lib.oracle_optimal_tile_size.argtypes = [ctypes.c_int, ctypes.c_int]
        self.lib.oracle_num_tiles.argtypes = [ctypes.c_int, ctypes.c_int]
        self.lib.oracle_reference_h2d_bytes.argtypes = [i64, i64, i64, ctypes.c_int, ctypes.c_int, i64, ctypes.c_int]
        self.lib.oracle_reference_h2d_bytes.restype = i64

    def gemm(self, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, wide=False):
        """In place on c (1-D column-major storage)."""
        dt = c.dtype
        al, be = np.array([alpha], dtype=dt), np.array([beta], dtype=dt)
        rc = self.lib.oracle_gemm_ex(CODES[dt], _c(ta), _c(tb), m, n, k, _vp(al), _vp(a), lda, _vp(b), ldb, _vp(be), _vp(c), ldc, 1 if wide else 0)
        if rc != 0:
            raise ValueError(f"oracle_gemm rc={rc}")
        return c

    def fixture_reset(self, seed=42):
        self.lib.oracle_fixture_reset(seed)

    def fixture_fill(self, arr):
        self.lib.oracle_fixture_fill(CODES[arr.dtype], _vp(arr), arr.size)
        return arr

    def fixture_abc(self, dtype, na, nb, nc):
        """Exactly what tests/test-multiply.cpp:269-274 does: reset generator, fill A, then B, then C."""
        self.fixture_reset(42)
        a, b, c = (np.empty(x, dtype=dtype) for x in (na, nb, nc))
        self.fixture_fill(a); self.fixture_fill(b); self.fixture_fill(c)
        return a, b, c


class Reference:
    """The UNMODIFIED reference library compiled by oracle/Makefile: `cpu` = over the CUDA/cuBLAS emulation
    (runs anywhere), else the real cuBLAS build (GPU box)."""

    def __init__(self, cpu: bool):
        path = ORACLE_DIR / "_ref" / ("libtiledmm_ref_cpu.so" if cpu else "libtiledmm_ref.so")
        if not path.exists():
            raise FileNotFoundError(path)
        self.cpu = cpu
        self.lib = ctypes.CDLL(str(path))
        L = self.lib
        L.ref_ctx_create.restype = ctypes.c_void_p
        L.ref_ctx_create.argtypes = [ctypes.c_int] * 5
        L.ref_ctx_destroy.argtypes = [ctypes.c_int, ctypes.c_void_p]
        L.ref_gemm.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_char, ctypes.c_char, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                               ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                               ctypes.c_int]
        L.ref_fetch_device_c.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]
        L.ref_optimal_tile_sizes.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        if cpu:
            for f in ("emul_h2d_bytes", "emul_d2h_bytes", "emul_gemm_calls"):
                getattr(L, f).restype = ctypes.c_uint64

    def context(self, dtype, streams=2, tm=5000, tn=5000, tk=5000):
        return RefContext(self, np.dtype(dtype), streams, tm, tn, tk)


class RefContext:
    def __init__(self, ref, dtype, streams, tm, tn, tk):
        self.ref, self.dtype, self.code = ref, dtype, CODES[dtype]
        self.h = ctypes.c_void_p(ref.lib.ref_ctx_create(self.code, streams, tm, tn, tk))
        assert self.h, "reference context creation failed"

    def gemm(self, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, pin=False, copy_c_back=True):
        al, be = np.array([alpha], dtype=self.dtype), np.array([beta], dtype=self.dtype)
        rc = self.ref.lib.ref_gemm(self.h, self.code, _c(ta), _c(tb), m, n, k, _vp(al), _vp(a), lda, _vp(b), ldb, _vp(be), _vp(c), ldc, int(pin), int(copy_c_back))
        assert rc == 0, "reference gemm raised"

    def fetch_device_c(self, count):
        out = np.empty(count, dtype=self.dtype)
        assert self.ref.lib.ref_fetch_device_c(self.h, self.code, _vp(out), count) == 0
        return out

    def optimal_tile_sizes(self, m, n, k):
        out = (ctypes.c_int * 3)()
        assert self.ref.lib.ref_optimal_tile_sizes(self.h, self.code, m, n, k, out) == 0
        return tuple(out)

    def close(self):
        if self.h:
            self.ref.lib.ref_ctx_destroy(self.code, self.h)
            self.h = None

    __del__ = close


def stored_shape(trans, rows_op, cols_op):
    """Stored (untransposed) shape of an operand whose op() is rows_op x cols_op (reference tiled_mm.cpp:507-511)."""
    return (rows_op, cols_op) if trans.upper() == "N" else (cols_op, rows_op)


def random_matrix(rng, dtype, count, lo=-1.0, hi=1.0):
    dt = np.dtype(dtype)
    if dt.kind == "c":
        real = np.float32 if dt == np.complex64 else np.float64
        return (rng.uniform(lo, hi, count).astype(real) + 1j * rng.uniform(lo, hi, count).astype(real)).astype(dt)
    return rng.uniform(lo, hi, count).astype(dt)


def rel_err(c, c_ref, k, amax, bmax, m, n, ldc):
    """north_star metric: max|C - C_ref| / (k * max|A| * max|B|), over the m x n window of an ld = ldc buffer."""
    c2 = c.reshape(n, ldc)[:, :m] if c.size == n * ldc else c
    r2 = c_ref.reshape(n, ldc)[:, :m] if c_ref.size == n * ldc else c_ref
    return float(np.max(np.abs(c2 - r2))) / (max(k, 1) * max(amax, 1e-300) * max(bmax, 1e-300))


# tolerance on that metric per dtype (FP64: north_star's 1e-15; FP32: true-FP32 accumulation, ~eps_f32 * sqrt-ish growth)
TOL = {np.dtype(np.float64): 1e-15, np.dtype(np.complex128): 2e-15, np.dtype(np.float32): 2e-6, np.dtype(np.complex64): 4e-6}


def window(c, m, n, ldc):
    return c.reshape(n, ldc)[:, :m]
