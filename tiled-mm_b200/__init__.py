"""tiled_mm_b200 — host-side Python mirror of the Tiled-MM public API over the C ABI.

Mirrors, name for name, the reference's C++ surface for the hot path
(reference src/Tiled-MM/tiled_mm.hpp:69-79, mm_handle.hpp:11-76, util.hpp:57-118):

    ctx = make_context(np.float64, streams, tile_m, tile_n, tile_k)     # gpu::make_context<double>(...)
    gemm(ctx, 'N', 'T', m, n, k, alpha, a, ld_a, b, ld_b, beta, c, ld_c,
         pin_host_buffers=False, copy_c_back=True)                      # gpu::gemm(*ctx, ...)
    a = malloc_pinned(np.float64, N, 1.0)                                # gpu::malloc_pinned<double>(N, 1)
    dc = ctx.get_full_device_buffer_c(); dc.data(); dc.size()            # ctx->get_full_device_buffer_c()
    copy_to_host(dc.data(), c_host, m * n)                               # gpu::copy_to_host(...)

Everything numeric happens in libtiledmm_b200.so (hand-written sm_100a kernels + scheduler).
There is NO CPU fallback: if the shared library is missing or no B200 is present, calls raise.
Errors mirror the reference: a failing call raises RuntimeError("GPU ERROR: ...") (util.hpp:13-27).
"""
from __future__ import annotations

import ctypes
import os
import re
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libtiledmm_b200.so"
HEADER_PATH = _HERE.parent / "include" / "tiled_mm_b200.h"

TMM_F32, TMM_F64, TMM_C32, TMM_C64 = 0, 1, 2, 3
TMM_OK, TMM_ERR_INVALID, TMM_ERR_CUDA, TMM_ERR_NOMEM, TMM_ERR_NOGPU = 0, -1, -2, -3, -4

_DTYPES = {
    np.dtype(np.float32): TMM_F32,
    np.dtype(np.float64): TMM_F64,
    np.dtype(np.complex64): TMM_C32,
    np.dtype(np.complex128): TMM_C64,
}
_NP_OF = {v: k for k, v in _DTYPES.items()}


class CallStats(ctypes.Structure):
    _fields_ = [
        ("h2d_bytes", ctypes.c_uint64),
        ("d2h_bytes", ctypes.c_uint64),
        ("kernel_launches", ctypes.c_uint64),
        ("h2d_copies", ctypes.c_uint64),
        ("d2h_copies", ctypes.c_uint64),
        ("wall_ms", ctypes.c_double),
        ("kernel_ms", ctypes.c_double),
        ("regime", ctypes.c_int),
        ("c_blocks", ctypes.c_int),
        ("k_chunks", ctypes.c_int),
        ("peer_bytes", ctypes.c_uint64),
    ]


_lib = None


def declared_symbols() -> list[str]:
    """Every entry point include/tiled_mm_b200.h declares."""
    text = HEADER_PATH.read_text()
    return sorted(set(re.findall(r"TMM_API[^;(]*?\b(tmm_\w+)\s*\(", text)))


def load_library() -> ctypes.CDLL:
    """Load the CUDA extension.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"tiled_mm_b200: CUDA extension {LIB_PATH} is missing - build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)"
        )
    lib = ctypes.CDLL(str(LIB_PATH))
    vp, i64, ci, cc, sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_char, ctypes.c_size_t
    lib.tmm_context_create.argtypes = [ci, ci, ci, ci, ci, ctypes.POINTER(vp)]
    lib.tmm_context_destroy.argtypes = [vp]
    lib.tmm_context_destroy.restype = None
    lib.tmm_gemm.argtypes = [vp, cc, cc, i64, i64, i64, vp, vp, i64, vp, i64, vp, vp, i64, ci, ci]
    lib.tmm_context_device_c.argtypes = [vp]
    lib.tmm_context_device_c.restype = vp
    lib.tmm_context_device_c_size.argtypes = [vp]
    lib.tmm_context_device_c_size.restype = sz
    lib.tmm_context_reserve_device_c.argtypes = [vp, i64, i64]
    lib.tmm_context_stream.argtypes = [vp, ci, ci]
    lib.tmm_context_stream.restype = vp
    lib.tmm_context_optimal_tile_sizes.argtypes = [vp, ci, ci, ci] + [ctypes.POINTER(ci)] * 3
    lib.tmm_context_get_max_tile_sizes.argtypes = [vp] + [ctypes.POINTER(ci)] * 3
    lib.tmm_context_get_num_streams.argtypes = [vp]
    lib.tmm_context_set_streams_and_tiles.argtypes = [vp, ci, ci, ci, ci]
    lib.tmm_context_dtype.argtypes = [vp]
    lib.tmm_malloc_pinned.argtypes = [sz, ctypes.POINTER(vp)]
    lib.tmm_free_pinned.argtypes = [vp]
    if hasattr(lib, "tmm_malloc_pinned_large"):
        lib.tmm_malloc_pinned_large.argtypes = [sz, ctypes.POINTER(vp)]
    lib.tmm_malloc_device.argtypes = [sz, ctypes.POINTER(vp)]
    lib.tmm_free_device.argtypes = [vp]
    lib.tmm_copy_to_device.argtypes = [vp, vp, sz]
    lib.tmm_copy_to_host.argtypes = [vp, vp, sz]
    lib.tmm_device_gemm.argtypes = [ci, cc, cc, i64, i64, i64, vp, vp, i64, vp, i64, vp, vp, i64, vp]
    lib.tmm_device_gemm_bf16.argtypes = [cc, cc, i64, i64, i64, ctypes.c_float, vp, i64, vp, i64, ctypes.c_float, vp, i64, vp]
    lib.tmm_context_last_stats.argtypes = [vp, ctypes.POINTER(CallStats)]
    lib.tmm_context_set_profiling.argtypes = [vp, ci]
    lib.tmm_context_set_device_budget.argtypes = [vp, sz]
    lib.tmm_total_kernel_launches.restype = ctypes.c_uint64
    lib.tmm_last_error.restype = ctypes.c_char_p
    lib.tmm_version.restype = ctypes.c_char_p
    lib.tmm_optimal_tile_size.argtypes = [ci, ci]
    lib.tmm_set_f32_math.argtypes = [ci]
    lib.tmm_set_c32_math.argtypes = [ci]
    lib.tmm_plan_describe.argtypes = [ci, cc, cc, i64, i64, i64, ci, ci, sz, ci, ci, ci, ci, ci, ctypes.c_char_p, sz]
    lib.tmm_grid_shape.argtypes = [ci, ctypes.POINTER(ci), ctypes.POINTER(ci)]
    lib.tmm_share_range.argtypes = [i64, ci, ci, ctypes.POINTER(i64), ctypes.POINTER(i64)]
    lib.tmm_dist_unique_id.argtypes = [vp]
    lib.tmm_context_attach_grid.argtypes = [vp, ci, ci, ci, ci, vp, vp]
    lib.tmm_context_grid.argtypes = [vp] + [ctypes.POINTER(ci)] * 4
    lib.tmm_context_set_devices.argtypes = [vp, ci, ctypes.POINTER(ci)]
    lib.tmm_context_num_devices.argtypes = [vp]
    lib.tmm_context_child.argtypes = [vp, ci]
    lib.tmm_context_child.restype = vp
    lib.tmm_memcpy_2d_async.argtypes = [vp, sz, vp, sz, sz, sz, ci, vp]
    if hasattr(lib, "tmm_probe_fp64_peak"):  # (the emulated-runtime build of tests/emul has no device code)
        lib.tmm_probe_fp64_peak.argtypes = [ctypes.POINTER(ctypes.c_double)]
        lib.tmm_probe_host_links.argtypes = [ci, ctypes.POINTER(ci), sz, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    _lib = lib
    return lib


def _check(rc: int) -> None:
    if rc != TMM_OK:
        msg = load_library().tmm_last_error().decode(errors="replace")
        if rc == TMM_ERR_INVALID:
            raise ValueError(f"tiled_mm_b200: {msg}")
        raise RuntimeError(f"GPU ERROR: {msg}" if not msg.startswith("GPU ERROR") else msg)


def dtype_code(dtype) -> int:
    try:
        return _DTYPES[np.dtype(dtype)]
    except KeyError:
        raise ValueError(f"unsupported scalar type {dtype}; Tiled-MM instantiates float, double, complex<float>, complex<double>") from None


def _ptr(x) -> ctypes.c_void_p:
    if x is None:
        return ctypes.c_void_p(0)
    if isinstance(x, (int, np.integer)):
        return ctypes.c_void_p(int(x))
    if isinstance(x, ctypes.c_void_p):
        return x
    if isinstance(x, np.ndarray):
        return ctypes.c_void_p(x.ctypes.data)
    if hasattr(x, "data_ptr"):  # torch tensor (host, pinned)
        return ctypes.c_void_p(x.data_ptr())
    raise TypeError(f"cannot take a pointer from {type(x)}")


def _scalar(dtype, v):
    return np.array([v], dtype=dtype)


class DeviceVector:
    """View of the context's full device C: the object ctx->get_full_device_buffer_c() returns
    (reference mm_handle.cpp:162-165, device_vector.hpp:67-80)."""

    def __init__(self, ctx: "MMHandle"):
        self._ctx = ctx

    def data(self) -> int:
        return load_library().tmm_context_device_c(self._ctx._h) or 0

    def size(self) -> int:
        return int(load_library().tmm_context_device_c_size(self._ctx._h))


class MMHandle:
    """gpu::mm_handle<Scalar> (reference mm_handle.hpp:11-58)."""

    def __init__(self, dtype, streams: int = 2, max_tile_m: int = 5000, max_tile_n: int = 5000, max_tile_k: int = 5000):
        lib = load_library()
        self.dtype = np.dtype(dtype)
        h = ctypes.c_void_p()
        _check(lib.tmm_context_create(dtype_code(dtype), streams, max_tile_m, max_tile_n, max_tile_k, ctypes.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            load_library().tmm_context_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def get_num_streams(self) -> int:
        return load_library().tmm_context_get_num_streams(self._h)

    def set_num_streams(self, streams: int) -> None:
        tm, tn, tk = self.get_max_tile_sizes()
        self.set_streams_and_tiles(streams, tm, tn, tk)

    def set_streams_and_tiles(self, streams: int, tile_m: int, tile_n: int, tile_k: int) -> None:
        _check(load_library().tmm_context_set_streams_and_tiles(self._h, streams, tile_m, tile_n, tile_k))

    def set_tile_sizes(self, tile_m: int, tile_n: int | None = None, tile_k: int | None = None) -> None:
        """mm_handle::set_tile_sizes(m, n, k) / set_tile_sizes(size) (mm_handle.cpp:57-71): staging hints, clamped to the maxima."""
        tn = tile_m if tile_n is None else tile_n
        tk = tile_m if tile_k is None else tile_k
        self.set_streams_and_tiles(self.get_num_streams(), tile_m, tn, tk)

    def get_max_tile_sizes(self):
        a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _check(load_library().tmm_context_get_max_tile_sizes(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return a.value, b.value, c.value

    def optimal_tile_sizes(self, m: int, n: int, k: int):
        a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _check(load_library().tmm_context_optimal_tile_sizes(self._h, m, n, k, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return a.value, b.value, c.value

    def get_full_device_buffer_c(self) -> DeviceVector:
        return DeviceVector(self)

    def set_full_sizes(self, m: int, n: int, k: int = 1) -> None:
        """mm_handle::set_full_sizes (mm_handle.cpp:73-80): size the full device C to m x n elements now."""
        _check(load_library().tmm_context_reserve_device_c(self._h, m, n))

    def stream(self, kind: int = 0, index: int = 0) -> int:
        """The context's own CUDA streams (gpu_context::get_stream / get_result_stream): kind STREAM_COMPUTE / STREAM_H2D / STREAM_D2H."""
        return load_library().tmm_context_stream(self._h, kind, index) or 0

    # multi-GPU (not in the reference, which drives one device)
    def set_devices(self, n_devices: int, device_ids=None) -> None:
        """One process, many GPUs: every later gemm(copy_c_back=True) splits C over `n_devices` child contexts."""
        ids = (ctypes.c_int * n_devices)(*device_ids) if device_ids is not None else None
        _check(load_library().tmm_context_set_devices(self._h, n_devices, ids))

    def num_devices(self) -> int:
        return load_library().tmm_context_num_devices(self._h)

    def attach_grid(self, grid_rows: int, grid_cols: int, my_row: int, my_col: int, row_id: bytes | None, col_id: bytes | None) -> None:
        """One process per GPU: join a grid_rows x grid_cols grid; gemm() then computes this rank's C block (see multi_gpu.py)."""
        rb = ctypes.create_string_buffer(row_id, 128) if row_id is not None else None
        cb = ctypes.create_string_buffer(col_id, 128) if col_id is not None else None
        _check(load_library().tmm_context_attach_grid(self._h, grid_rows, grid_cols, my_row, my_col,
                                                      ctypes.cast(rb, ctypes.c_void_p) if rb is not None else None,
                                                      ctypes.cast(cb, ctypes.c_void_p) if cb is not None else None))

    def grid(self):
        v = [ctypes.c_int() for _ in range(4)]
        _check(load_library().tmm_context_grid(self._h, *(ctypes.byref(x) for x in v)))
        return tuple(x.value for x in v)

    # introspection (not in the reference)
    def last_stats(self) -> CallStats:
        st = CallStats()
        _check(load_library().tmm_context_last_stats(self._h, ctypes.byref(st)))
        return st

    def set_profiling(self, on: bool) -> None:
        _check(load_library().tmm_context_set_profiling(self._h, 1 if on else 0))

    def set_device_budget(self, nbytes: int) -> None:
        _check(load_library().tmm_context_set_device_budget(self._h, nbytes))


def make_context(dtype=np.float64, streams: int = 2, max_tile_m: int = 5000, max_tile_n: int = 5000, max_tile_k: int = 5000) -> MMHandle:
    """gpu::make_context<Scalar>(streams, max_tile_m, max_tile_n, max_tile_k); defaults 2 / 5000^3 (mm_handle.hpp:60-76)."""
    return MMHandle(dtype, streams, max_tile_m, max_tile_n, max_tile_k)


def gemm(handle: MMHandle, trans_a: str, trans_b: str, m: int, n: int, k: int, alpha, a, ld_a: int, b, ld_b: int, beta, c, ld_c: int,
         pin_host_buffers: bool = True, copy_c_back: bool = True) -> None:
    """gpu::gemm<Scalar>(handle, trans_a, trans_b, m, n, k, alpha, a, ld_a, b, ld_b, beta, c, ld_c, pin_host_buffers, copy_c_back)
    (reference tiled_mm.hpp:69-79).  a, b, c: host buffers (numpy arrays / pinned arrays / raw addresses), column-major."""
    lib = load_library()
    al, be = _scalar(handle.dtype, alpha), _scalar(handle.dtype, beta)
    for name, x in (("a", a), ("b", b), ("c", c)):
        if isinstance(x, np.ndarray) and x.dtype != handle.dtype:
            raise ValueError(f"{name} has dtype {x.dtype}, context is {handle.dtype}")
    ta = ctypes.c_char(trans_a.encode()[:1])
    tb = ctypes.c_char(trans_b.encode()[:1])
    _check(lib.tmm_gemm(handle._h, ta, tb, m, n, k, _ptr(al), _ptr(a), ld_a, _ptr(b), ld_b, _ptr(be), _ptr(c), ld_c,
                        1 if pin_host_buffers else 0, 1 if copy_c_back else 0))


class _PinnedOwner:
    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        try:
            if self.ptr and _lib is not None:
                _lib.tmm_free_pinned(ctypes.c_void_p(self.ptr))
        except Exception:
            pass


class _PinnedArray(np.ndarray):
    """ndarray over a cudaHostAlloc block; every view keeps the block alive through `_owner`."""
    _owner = None

    def __array_finalize__(self, obj):
        if obj is not None:
            self._owner = getattr(obj, "_owner", None)


def malloc_pinned(dtype, count: int, value=0) -> np.ndarray:
    """gpu::malloc_pinned<T>(N, value): cudaHostAlloc(flags 0) + fill (reference util.hpp:65-72).
    Returns a 1-D numpy array over the pinned block; the block is freed when the last view is collected."""
    lib = load_library()
    dt = np.dtype(dtype)
    p = ctypes.c_void_p()
    nbytes = max(1, count) * dt.itemsize
    _check(lib.tmm_malloc_pinned(nbytes, ctypes.byref(p)))
    owner = _PinnedOwner(p.value)
    buf = (ctypes.c_byte * nbytes).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dt, count=count).view(_PinnedArray)
    arr._owner = owner
    arr.fill(value)
    return arr


def malloc_pinned_large(dtype, count: int) -> np.ndarray:
    """tmm_malloc_pinned_large: zero-filled pinned memory on 2 MiB pages, first touched by all cores and registered in one cudaHostRegister -
    an order of magnitude faster than cudaHostAlloc for the hundreds of GB the out-of-core configs need.  Same array type and lifetime rules
    as malloc_pinned (tmm_free_pinned knows both kinds)."""
    lib = load_library()
    dt = np.dtype(dtype)
    p = ctypes.c_void_p()
    nbytes = max(1, count) * dt.itemsize
    _check(lib.tmm_malloc_pinned_large(nbytes, ctypes.byref(p)))
    owner = _PinnedOwner(p.value)
    buf = (ctypes.c_byte * nbytes).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dt, count=count).view(_PinnedArray)
    arr._owner = owner
    return arr


def malloc_device(nbytes: int) -> int:
    p = ctypes.c_void_p()
    _check(load_library().tmm_malloc_device(nbytes, ctypes.byref(p)))
    return p.value


def free_device(ptr: int) -> None:
    _check(load_library().tmm_free_device(ctypes.c_void_p(ptr)))


def copy_to_device(host: np.ndarray, device_ptr: int, count: int | None = None) -> None:
    """gpu::copy_to_device(from, to, n) (util.hpp:79-82); count in elements."""
    n = host.size if count is None else count
    _check(load_library().tmm_copy_to_device(_ptr(host), ctypes.c_void_p(device_ptr), n * host.dtype.itemsize))


def copy_to_host(device_ptr: int, host: np.ndarray, count: int | None = None) -> None:
    """gpu::copy_to_host(from, to, n) (util.hpp:85-88); count in elements."""
    n = host.size if count is None else count
    _check(load_library().tmm_copy_to_host(ctypes.c_void_p(device_ptr), _ptr(host), n * host.dtype.itemsize))


def device_gemm(dtype, trans_a: str, trans_b: str, m: int, n: int, k: int, alpha, a_dev: int, ld_a: int, b_dev: int, ld_b: int, beta, c_dev: int,
                ld_c: int, stream: int = 0) -> None:
    """blas_api::{s,d,c,z}gemm replacement on device pointers (gpu_blas_api.hpp:194-252)."""
    dt = np.dtype(dtype)
    al, be = _scalar(dt, alpha), _scalar(dt, beta)
    _check(load_library().tmm_device_gemm(dtype_code(dt), ctypes.c_char(trans_a.encode()[:1]), ctypes.c_char(trans_b.encode()[:1]), m, n, k,
                                          _ptr(al), ctypes.c_void_p(a_dev), ld_a, ctypes.c_void_p(b_dev), ld_b, _ptr(be), ctypes.c_void_p(c_dev), ld_c,
                                          ctypes.c_void_p(stream)))


def device_gemm_bf16(trans_a: str, trans_b: str, m: int, n: int, k: int, alpha: float, a_dev: int, ld_a: int, b_dev: int, ld_b: int, beta: float,
                     c_dev: int, ld_c: int, stream: int = 0) -> None:
    """C (float32) = alpha op(A) op(B) + beta C with A, B stored as bfloat16 on the device (additive entry point, tmm_device_gemm_bf16)."""
    _check(load_library().tmm_device_gemm_bf16(ctypes.c_char(trans_a.encode()[:1]), ctypes.c_char(trans_b.encode()[:1]), m, n, k, alpha,
                                               ctypes.c_void_p(a_dev), ld_a, ctypes.c_void_p(b_dev), ld_b, beta, ctypes.c_void_p(c_dev), ld_c,
                                               ctypes.c_void_p(stream)))


def optimal_tile_size(dim: int, max_tile: int) -> int:
    """Pure host logic of mm_handle::optimal_tile_sizes for one dimension (reference mm_handle.cpp:89-110)."""
    return load_library().tmm_optimal_tile_size(dim, max_tile)


def plan_describe(dtype, trans_a: str, trans_b: str, m: int, n: int, k: int, beta_nonzero: bool, copy_c_back: bool, budget_bytes: int,
                  streams: int = 2, tile_m: int = 5000, tile_n: int = 5000, tile_k: int = 5000, sm_count: int = 148) -> dict:
    """The scheduler's plan for a call as JSON (pure host logic, needs no GPU)."""
    import json
    buf = ctypes.create_string_buffer(1 << 20)
    rc = load_library().tmm_plan_describe(dtype_code(dtype), ctypes.c_char(trans_a.encode()[:1]), ctypes.c_char(trans_b.encode()[:1]), m, n, k,
                                          1 if beta_nonzero else 0, 1 if copy_c_back else 0, budget_bytes, streams, tile_m, tile_n, tile_k, sm_count,
                                          buf, len(buf))
    _check(rc)
    return json.loads(buf.value.decode())


def grid_shape(n_gpus: int):
    """(grid_rows, grid_cols) of the C-block grid for n GPUs: 1->1x1, 2->1x2, 4->2x2, 8->2x4."""
    a, b = ctypes.c_int(), ctypes.c_int()
    _check(load_library().tmm_grid_shape(n_gpus, ctypes.byref(a), ctypes.byref(b)))
    return a.value, b.value


def share_range(extent: int, parts: int, index: int):
    """[lo, hi) of share `index` of a balanced split of `extent` into `parts` (C blocks and panel upload shares)."""
    lo, hi = ctypes.c_int64(), ctypes.c_int64()
    _check(load_library().tmm_share_range(extent, parts, index, ctypes.byref(lo), ctypes.byref(hi)))
    return lo.value, hi.value


def dist_unique_id() -> bytes:
    buf = ctypes.create_string_buffer(128)
    _check(load_library().tmm_dist_unique_id(ctypes.cast(buf, ctypes.c_void_p)))
    return buf.raw


def device_count() -> int:
    return load_library().tmm_device_count()


def probe_fp64_peak() -> float:
    """FP64 tensor (DMMA.8x8x4) issue rate of the current device in TFLOP/s: the ceiling the DGEMM / ZGEMM kernels are measured against."""
    v = ctypes.c_double(0.0)
    _check(load_library().tmm_probe_fp64_peak(ctypes.byref(v)))
    return float(v.value)


def probe_host_links(device_ids=None, n_devices: int | None = None, nbytes: int = 64 << 20):
    """[(h2d GB/s, d2h GB/s)] per device with ALL listed devices copying both ways at once (default: every device of the box)."""
    lib = load_library()
    ids = list(device_ids) if device_ids is not None else list(range(n_devices if n_devices is not None else device_count()))
    arr = (ctypes.c_int * len(ids))(*ids)
    up = (ctypes.c_double * len(ids))()
    down = (ctypes.c_double * len(ids))()
    _check(lib.tmm_probe_host_links(len(ids), arr, nbytes, up, down))
    return [(float(up[i]), float(down[i])) for i in range(len(ids))]


def total_kernel_launches() -> int:
    return int(load_library().tmm_total_kernel_launches())


MATH_SIMT, MATH_TF32, MATH_FP32 = 0, 1, 3
STREAM_COMPUTE, STREAM_H2D, STREAM_D2H = 0, 1, 2


def set_f32_math(mode: int) -> None:
    """Math mode of the float GEMM (tmm_set_f32_math): MATH_FP32 (default: FP32-accurate 3xTF32 on tcgen05), MATH_TF32, MATH_SIMT."""
    _check(load_library().tmm_set_f32_math(int(mode)))


CMATH_SIMT, CMATH_TC = 0, 3


def set_c32_math(mode: int) -> None:
    """Math mode of the complex<float> GEMM (tmm_set_c32_math): CMATH_SIMT (default) or CMATH_TC (real embedding on tcgen05, opt-in)."""
    _check(load_library().tmm_set_c32_math(int(mode)))


def get_c32_math() -> int:
    return int(load_library().tmm_get_c32_math())


def get_f32_math() -> int:
    return int(load_library().tmm_get_f32_math())


def version() -> str:
    return load_library().tmm_version().decode()
