// Shared plumbing of the device GEMM layer: dtype dispatch, C scaling, launch accounting,
// driver entry point lookup.
#include "tmm_blas.h"

#include <algorithm>

#include <atomic>
#include <mutex>
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cuComplex.h>

namespace tmm {

static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
uint64_t launch_count() { return g_launches.load(std::memory_order_relaxed); }

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!cached[dev]) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cached[dev] = n > 0 ? n : 1;
    }
    return cached[dev];
}

void* tensormap_encode_fn() {
    // resolved once, thread-safe: the multi-GPU path launches from one host thread per device
    static void* fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        cudaDriverEntryPointQueryResult qres;
        void* p = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = p;
    });
    return fn;
}

// ---- stream-ordered scratch ----------------------------------------------------------------------
// Operand preparation (complex embedding, BF16 widening, re-pitching, int8 slices) works in stream-ordered scratch.  The device's DEFAULT pool
// releases its free blocks to the driver at every synchronisation (release threshold 0): each host-to-host call ends in one, so the next call's
// first launches would pay a physical allocation of up to a few GB again, and the pool's attributes belong to the application anyway.  The
// library therefore keeps one pool of its own per device that holds on to what it has grown to; scratch_trim() hands the free blocks back
// (the scheduler calls it when it re-reads the free device memory, so cached scratch never shrinks an out-of-core plan).
// TMM_POOL_KEEP=0: the default pool with its default behaviour (A/B switch).
namespace {
struct ScratchPools {
    std::mutex mu;
    cudaMemPool_t pool[64] = {};
    bool tried[64] = {};
};
ScratchPools& pools() { static ScratchPools* p = new ScratchPools; return *p; }  // leaked on purpose: no teardown order problem at exit
cudaMemPool_t scratch_pool() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    ScratchPools& sp = pools();
    std::lock_guard<std::mutex> lk(sp.mu);
    if (!sp.tried[dev]) {
        sp.tried[dev] = true;
        const char* v = getenv("TMM_POOL_KEEP");
        if (!(v && v[0] == '0')) {
            cudaMemPoolProps props = {};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            cudaMemPool_t mp = nullptr;
            if (cudaMemPoolCreate(&mp, &props) == cudaSuccess) {
                uint64_t keep = UINT64_MAX;
                cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep);
                sp.pool[dev] = mp;
            }
            cudaGetLastError();
        }
    }
    return sp.pool[dev];
}
}  // namespace

cudaError_t scratch_alloc(void** p, size_t bytes, cudaStream_t st) {
    cudaMemPool_t mp = scratch_pool();
    return mp ? cudaMallocFromPoolAsync(p, bytes, mp, st) : cudaMallocAsync(p, bytes, st);
}
void scratch_free(void* p, cudaStream_t st) { if (p) cudaFreeAsync(p, st); }
void scratch_trim() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
    ScratchPools& sp = pools();
    std::lock_guard<std::mutex> lk(sp.mu);
    if (sp.pool[dev]) cudaMemPoolTrimTo(sp.pool[dev], 0);
}

// ---- float GEMM math mode and dispatch ---------------------------------------------------------
static std::atomic<int> g_f32_mode{-1};
int f32_math_mode() {
    int v = g_f32_mode.load(std::memory_order_relaxed);
    if (v < 0) {
        const char* e = getenv("TMM_F32_MATH");  // "fp32" (default) | "tf32" | "simt"
        v = 3;
        if (e && (!strcmp(e, "tf32") || !strcmp(e, "TF32"))) v = 1;
        else if (e && (!strcmp(e, "simt") || !strcmp(e, "SIMT"))) v = 0;
        g_f32_mode.store(v, std::memory_order_relaxed);
    }
    return v;
}
void set_f32_math_mode(int mode) { g_f32_mode.store(mode == 1 ? 1 : (mode == 0 ? 0 : 3), std::memory_order_relaxed); }

static cudaError_t sgemm_repitched(char ta, char tb, int m, int n, int k, float alpha, const float* a, int64_t lda, const float* b, int64_t ldb, float beta,
                                   float* c, int64_t ldc, cudaStream_t st, int mode);

cudaError_t sgemm_launch(char ta, char tb, int m, int n, int k, float alpha, const float* a, int64_t lda, const float* b, int64_t ldb, float beta,
                         float* c, int64_t ldc, cudaStream_t st) {
    const int mode = f32_math_mode();
    if (mode != 0 && sgemm_tc_eligible(a, lda, b, ldb)) return sgemm_tc_launch(ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, st, mode);
    if (mode != 0) {
        // outside the TMA contract (base not 16-byte aligned, ld not a multiple of 4 floats): re-pitch once with the copy engine and stay on the
        // tensor cores, as the FP64 path does (the SIMT kernel is ~4x slower than what the reference's cuBLAS call delivers here)
        cudaError_t e = sgemm_repitched(ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, st, mode);
        if (e != cudaErrorMemoryAllocation) return e;
        cudaGetLastError();
    }
    return sgemm_simt_launch(ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, st);
}

// ---- complex<float> GEMM math mode and dispatch ------------------------------------------------
static std::atomic<int> g_c32_mode{-1};
int c32_math_mode() {
    int v = g_c32_mode.load(std::memory_order_relaxed);
    if (v < 0) {
        const char* e = getenv("TMM_C32_MATH");  // "tc" (default since round 2: 128 TF vs 45 TF SIMT vs 71 TF cuBLAS CGEMM at 8192^3) | "simt"
        v = (e && (!strcmp(e, "simt") || !strcmp(e, "SIMT"))) ? 0 : 3;
        g_c32_mode.store(v, std::memory_order_relaxed);
    }
    return v;
}
void set_c32_math_mode(int mode) { g_c32_mode.store(mode == 3 ? 3 : 0, std::memory_order_relaxed); }

cudaError_t cgemm_launch(char ta, char tb, int m, int n, int k, const float* al, const void* a, int64_t lda, const void* b, int64_t ldb,
                         const float* be, void* c, int64_t ldc, cudaStream_t st) {
    if (c32_math_mode() == 3 && f32_math_mode() != 0 && (reinterpret_cast<uintptr_t>(a) & 7) == 0 && (reinterpret_cast<uintptr_t>(b) & 7) == 0) {
        cudaError_t e = cgemm_tc_launch(ta, tb, m, n, k, al, a, lda, b, ldb, be, c, ldc, st);
        if (e != cudaErrorMemoryAllocation) return e;  // no scratch: the SIMT kernel needs none
    }
    return cgemm_simt_launch(ta, tb, m, n, k, al, a, lda, b, ldb, be, c, ldc, st);
}

// ---- C = beta * C --------------------------------------------------------------------------
template <typename T> struct Ops;
template <> struct Ops<float> {
    static __device__ float zero() { return 0.f; }
    static __device__ float mul(float a, float b) { return a * b; }
};
template <> struct Ops<double> {
    static __device__ double zero() { return 0.0; }
    static __device__ double mul(double a, double b) { return a * b; }
};
template <> struct Ops<cuFloatComplex> {
    static __device__ cuFloatComplex zero() { return make_cuFloatComplex(0.f, 0.f); }
    static __device__ cuFloatComplex mul(cuFloatComplex a, cuFloatComplex b) { return cuCmulf(a, b); }
};
template <> struct Ops<cuDoubleComplex> {
    static __device__ cuDoubleComplex zero() { return make_cuDoubleComplex(0.0, 0.0); }
    static __device__ cuDoubleComplex mul(cuDoubleComplex a, cuDoubleComplex b) { return cuCmul(a, b); }
};

template <typename T>
__global__ void scale_kernel(T* c, int64_t ldc, int64_t m, int64_t n, T beta, int zero) {
    const int64_t total = m * n;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t col = idx / m, row = idx - col * m;
        T* p = c + col * ldc + row;
        *p = zero ? Ops<T>::zero() : Ops<T>::mul(beta, *p);
    }
}

template <typename T>
static cudaError_t scale_impl(int64_t m, int64_t n, T beta, bool zero, void* c, int64_t ldc, cudaStream_t st) {
    if (m <= 0 || n <= 0) return cudaSuccess;
    int64_t total = m * n;
    int blocks = (int)((total + 255) / 256 < (int64_t)sm_count() * 8 ? (total + 255) / 256 : (int64_t)sm_count() * 8);
    scale_kernel<T><<<blocks, 256, 0, st>>>(static_cast<T*>(c), ldc, m, n, beta, zero ? 1 : 0);
    count_launch();
    return cudaGetLastError();
}

cudaError_t device_scale(int dtype, int64_t m, int64_t n, const void* beta, void* c, int64_t ldc, cudaStream_t st) {
    switch (dtype) {
    case F32: { float b = *static_cast<const float*>(beta); return scale_impl<float>(m, n, b, b == 0.f, c, ldc, st); }
    case F64: { double b = *static_cast<const double*>(beta); return scale_impl<double>(m, n, b, b == 0.0, c, ldc, st); }
    case C32: { const float* b = static_cast<const float*>(beta); return scale_impl<cuFloatComplex>(m, n, make_cuFloatComplex(b[0], b[1]), b[0] == 0.f && b[1] == 0.f, c, ldc, st); }
    case C64: { const double* b = static_cast<const double*>(beta); return scale_impl<cuDoubleComplex>(m, n, make_cuDoubleComplex(b[0], b[1]), b[0] == 0.0 && b[1] == 0.0, c, ldc, st); }
    }
    return cudaErrorInvalidValue;
}

// ---- C += beta * S  (the scheduler adds the caller's beta * C at the END of a block's accumulation: its upload then no longer gates the first GEMM)
template <typename T> struct Fma;
template <> struct Fma<float> { static __device__ float f(float b, float s, float c) { return c + b * s; } };
template <> struct Fma<double> { static __device__ double f(double b, double s, double c) { return c + b * s; } };
template <> struct Fma<cuFloatComplex> { static __device__ cuFloatComplex f(cuFloatComplex b, cuFloatComplex s, cuFloatComplex c) { return cuCaddf(c, cuCmulf(b, s)); } };
template <> struct Fma<cuDoubleComplex> { static __device__ cuDoubleComplex f(cuDoubleComplex b, cuDoubleComplex s, cuDoubleComplex c) { return cuCadd(c, cuCmul(b, s)); } };

template <typename T>
__global__ void add_scaled_kernel(T* c, int64_t ldc, const T* __restrict__ s, int64_t lds, int64_t m, int64_t n, T beta) {
    const int64_t total = m * n;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t col = idx / m, row = idx - col * m;
        T* p = c + col * ldc + row;
        *p = Fma<T>::f(beta, s[col * lds + row], *p);
    }
}

template <typename T>
static cudaError_t add_scaled_impl(int64_t m, int64_t n, T beta, const void* s, int64_t lds, void* c, int64_t ldc, cudaStream_t st) {
    if (m <= 0 || n <= 0) return cudaSuccess;
    const int64_t total = m * n;
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 8);
    add_scaled_kernel<T><<<blocks, 256, 0, st>>>(static_cast<T*>(c), ldc, static_cast<const T*>(s), lds, m, n, beta);
    count_launch();
    return cudaGetLastError();
}

cudaError_t device_add_scaled(int dtype, int64_t m, int64_t n, const void* beta, const void* s, int64_t lds, void* c, int64_t ldc, cudaStream_t st) {
    switch (dtype) {
    case F32: return add_scaled_impl<float>(m, n, *static_cast<const float*>(beta), s, lds, c, ldc, st);
    case F64: return add_scaled_impl<double>(m, n, *static_cast<const double*>(beta), s, lds, c, ldc, st);
    case C32: { const float* b = static_cast<const float*>(beta); return add_scaled_impl<cuFloatComplex>(m, n, make_cuFloatComplex(b[0], b[1]), s, lds, c, ldc, st); }
    case C64: { const double* b = static_cast<const double*>(beta); return add_scaled_impl<cuDoubleComplex>(m, n, make_cuDoubleComplex(b[0], b[1]), s, lds, c, ldc, st); }
    }
    return cudaErrorInvalidValue;
}

// ---- operands that do not meet the TMA contract (FP64 paths) --------------------------------------------------------
// cuBLAS takes any pointer / leading dimension; the TMA-fed DMMA kernels need a 16-byte aligned base and a pitch that is a
// multiple of 16 bytes.  The scheduler's own panels always qualify (it picks the device pitch), but blas_api::dgemm is also a
// public entry point (reference tests/test-multiply.cpp:44 calls it with ld = k = 1357).  Such an operand is re-pitched once
// by the copy engine into a stream-ordered scratch block (cudaMallocAsync -> 2-D D2D copy -> kernel -> cudaFreeAsync), so the
// arithmetic stays on the tensor pipe; if the pool allocation fails the FP64 SIMT kernel takes the call.
namespace {
struct Operand {
    const void* p;
    int64_t ld;
    void* scratch = nullptr;
};
bool tma_ok(const void* p, int64_t ld, size_t es) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (((size_t)ld * es) & 15) == 0; }
cudaError_t repitch(Operand& op, int64_t rows, int64_t cols, size_t es, cudaStream_t st) {
    if (tma_ok(op.p, op.ld, es) || rows <= 0 || cols <= 0) return cudaSuccess;
    const int64_t q = 128 / (int64_t)es, pitch = (rows + q - 1) / q * q;
    void* buf = nullptr;
    cudaError_t e = scratch_alloc(&buf, (size_t)pitch * (size_t)cols * es, st);
    if (e != cudaSuccess) return e;
    e = cudaMemcpy2DAsync(buf, (size_t)pitch * es, op.p, (size_t)op.ld * es, (size_t)rows * es, (size_t)cols, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) { scratch_free(buf, st); return e; }
    op.p = buf; op.ld = pitch; op.scratch = buf;
    return cudaSuccess;
}
}  // namespace

static cudaError_t sgemm_repitched(char ta, char tb, int m, int n, int k, float alpha, const float* a, int64_t lda, const float* b, int64_t ldb, float beta,
                                   float* c, int64_t ldc, cudaStream_t st, int mode) {
    Operand A{a, lda}, B{b, ldb};
    cudaError_t e = repitch(A, ta == 'N' ? m : k, ta == 'N' ? k : m, sizeof(float), st);
    if (e == cudaSuccess) e = repitch(B, tb == 'N' ? k : n, tb == 'N' ? n : k, sizeof(float), st);
    if (e == cudaSuccess) e = sgemm_tc_launch(ta, tb, m, n, k, alpha, static_cast<const float*>(A.p), A.ld, static_cast<const float*>(B.p), B.ld, beta, c, ldc, st, mode);
    else e = cudaErrorMemoryAllocation;  // no scratch: the caller falls back to the SIMT kernel on the operands as they are
    scratch_free(A.scratch, st);
    scratch_free(B.scratch, st);
    return e;
}

static cudaError_t fp64_gemm(int dtype, char ta, char tb, int m, int n, int k, const void* alpha, const void* a, int64_t lda, const void* b, int64_t ldb,
                             const void* beta, void* c, int64_t ldc, cudaStream_t st) {
    const size_t es = dtype_size(dtype);
    if (dtype == F64 && f64_i8_slices() > 0) {  // experimental opt-in: FP64-accurate product on the int8 tensor cores
        cudaError_t ei = dgemm_i8_launch(ta, tb, m, n, k, *static_cast<const double*>(alpha), static_cast<const double*>(a), lda, static_cast<const double*>(b), ldb,
                                         *static_cast<const double*>(beta), static_cast<double*>(c), ldc, st, f64_i8_slices());
        if (ei != cudaErrorMemoryAllocation) return ei;  // no scratch: the DMMA kernel below needs none
    }
    Operand A{a, lda}, B{b, ldb};
    cudaError_t e = repitch(A, ta == 'N' ? m : k, ta == 'N' ? k : m, es, st);
    if (e == cudaSuccess) e = repitch(B, tb == 'N' ? k : n, tb == 'N' ? n : k, es, st);
    if (e == cudaSuccess) {
        e = dtype == F64 ? dgemm_launch(ta, tb, m, n, k, *static_cast<const double*>(alpha), static_cast<const double*>(A.p), A.ld,
                                        static_cast<const double*>(B.p), B.ld, *static_cast<const double*>(beta), static_cast<double*>(c), ldc, st)
                         : zgemm_launch(ta, tb, m, n, k, static_cast<const double*>(alpha), A.p, A.ld, B.p, B.ld, static_cast<const double*>(beta), c, ldc, st);
    } else {
        cudaGetLastError();  // no scratch: true-FP64 SIMT kernel on the operands as they are
        e = dtype == F64 ? dgemm_simt_launch(ta, tb, m, n, k, *static_cast<const double*>(alpha), static_cast<const double*>(a), lda,
                                             static_cast<const double*>(b), ldb, *static_cast<const double*>(beta), static_cast<double*>(c), ldc, st)
                         : zgemm_simt_launch(ta, tb, m, n, k, static_cast<const double*>(alpha), a, lda, b, ldb, static_cast<const double*>(beta), c, ldc, st);
    }
    scratch_free(A.scratch, st);
    scratch_free(B.scratch, st);
    return e;
}

cudaError_t device_gemm(int dtype, char trans_a, char trans_b, int64_t m, int64_t n, int64_t k, const void* alpha, const void* a, int64_t lda,
                        const void* b, int64_t ldb, const void* beta, void* c, int64_t ldc, cudaStream_t st) {
    char ta = (char)std::toupper((unsigned char)trans_a), tb = (char)std::toupper((unsigned char)trans_b);
    if ((ta != 'N' && ta != 'T' && ta != 'C') || (tb != 'N' && tb != 'T' && tb != 'C')) return cudaErrorInvalidValue;
    if (m < 0 || n < 0 || k < 0 || m > INT32_MAX || n > INT32_MAX || k > INT32_MAX) return cudaErrorInvalidValue;
    if (m == 0 || n == 0) return cudaSuccess;
    bool alpha_zero = false;
    switch (dtype) {
    case F32: alpha_zero = *static_cast<const float*>(alpha) == 0.f; break;
    case F64: alpha_zero = *static_cast<const double*>(alpha) == 0.0; break;
    case C32: alpha_zero = static_cast<const float*>(alpha)[0] == 0.f && static_cast<const float*>(alpha)[1] == 0.f; break;
    case C64: alpha_zero = static_cast<const double*>(alpha)[0] == 0.0 && static_cast<const double*>(alpha)[1] == 0.0; break;
    default: return cudaErrorInvalidValue;
    }
    if (k == 0 || alpha_zero) return device_scale(dtype, m, n, beta, c, ldc, st);  // BLAS convention
    switch (dtype) {
    case F32: return sgemm_launch(ta, tb, (int)m, (int)n, (int)k, *static_cast<const float*>(alpha), static_cast<const float*>(a), lda,
                                  static_cast<const float*>(b), ldb, *static_cast<const float*>(beta), static_cast<float*>(c), ldc, st);
    case F64: return fp64_gemm(F64, ta, tb, (int)m, (int)n, (int)k, alpha, a, lda, b, ldb, beta, c, ldc, st);
    case C32: return cgemm_launch(ta, tb, (int)m, (int)n, (int)k, static_cast<const float*>(alpha), a, lda, b, ldb, static_cast<const float*>(beta), c, ldc, st);
    case C64: return fp64_gemm(C64, ta, tb, (int)m, (int)n, (int)k, alpha, a, lda, b, ldb, beta, c, ldc, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace tmm
