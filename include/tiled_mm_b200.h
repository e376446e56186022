/*
 * tiled_mm_b200 — C ABI of the B200-native out-of-core GEMM (drop-in for the eth-cscs/Tiled-MM hot path).
 *
 * The reference has no C ABI: its boundary is the C++ template API in namespace gpu
 * (reference src/Tiled-MM/tiled_mm.hpp:62-80, mm_handle.hpp:11-76, util.hpp:57-118).  The C++
 * drop-in headers under include/Tiled-MM/ are thin wrappers over the entry points below, and the
 * same entry points are what a ctypes / cgo / JNI binding would use.
 *
 * Conventions: plain pointers and sizes only; dtype codes TMM_F32..TMM_C64; scalars (alpha, beta)
 * are passed by pointer to one element of the dtype (complex = {re, im}); all matrices are
 * column-major; every function returns TMM_OK (0) or a negative TMM_ERR_* code, and
 * tmm_last_error() returns a thread-local message for the last failure.  Like the reference
 * (util.hpp:13-27: message on stderr + std::runtime_error("GPU ERROR")), CUDA failures also print
 * the CUDA error string to stderr; the C++ wrappers turn a negative code into that exception.
 */
#ifndef TILED_MM_B200_H
#define TILED_MM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define TMM_API __attribute__((visibility("default")))
#else
#define TMM_API
#endif

#define TMM_F32 0 /* float                 reference instantiation tiled_mm.cpp:626-635 */
#define TMM_F64 1 /* double                tiled_mm.cpp:637-646 */
#define TMM_C32 2 /* std::complex<float>   tiled_mm.cpp:648-657 */
#define TMM_C64 3 /* std::complex<double>  tiled_mm.cpp:659-668 */

#define TMM_OK 0
#define TMM_ERR_INVALID (-1) /* bad argument (dtype, trans, negative size, ld too small) */
#define TMM_ERR_CUDA (-2)    /* a CUDA runtime/driver call failed ("GPU ERROR" in the reference) */
#define TMM_ERR_NOMEM (-3)   /* device or host allocation failed */
#define TMM_ERR_NOGPU (-4)   /* no CUDA device / kernels not loadable: there is NO CPU fallback */

typedef struct tmm_context tmm_context; /* opaque; replaces gpu::mm_handle<Scalar> (mm_handle.hpp:11-58) */

/* gpu::make_context<Scalar>(streams, max_tile_m, max_tile_n, max_tile_k)  — mm_handle.hpp:60-76, mm_handle.cpp:10-29.
 * Binds to the calling thread's current CUDA device (the reference never calls cudaSetDevice either,
 * mm_handle.cpp:18-24).  n_streams / tile sizes are hints: they bound staging granularity, never results. */
TMM_API int tmm_context_create(int dtype, int n_streams, int max_tile_m, int max_tile_n, int max_tile_k, tmm_context** out);
/* ~mm_handle()  — mm_handle.cpp:31-34 */
TMM_API void tmm_context_destroy(tmm_context* ctx);

/* gpu::gemm<Scalar>(handle, trans_a, trans_b, m, n, k, alpha, a, ld_a, b, ld_b, beta, c, ld_c,
 *                   pin_host_buffers, copy_c_back)  — tiled_mm.hpp:69-79, tiled_mm.cpp:492-624.
 * a, b, c are HOST pointers (device pointers are accepted as an extension: all operands on the context's device -> one launch on them,
 * result in c or in the context's device C; a mix of host and device operands is staged with direction-inferring copies).  Synchronous: returns after all device work; with copy_c_back != 0 host C
 * holds the result, otherwise it stays in the context's device C (tmm_context_device_c), column-major
 * with ld = m.  beta == 0 => C is never read (NaN-safe).  64-bit sizes: the reference's int offsets
 * overflow at 2^31 elements (tiled_matrix.cpp:62-67); these do not. */
TMM_API int tmm_gemm(tmm_context* ctx, char trans_a, char trans_b, int64_t m, int64_t n, int64_t k, const void* alpha, const void* a, int64_t ld_a,
             const void* b, int64_t ld_b, const void* beta, void* c, int64_t ld_c, int pin_host_buffers, int copy_c_back);

/* handle.get_full_device_buffer_c().data() / .size()  — mm_handle.cpp:162-165, device_vector.hpp:67-75.
 * Borrowed device pointer, valid until a later call grows it or the context is destroyed. */
TMM_API void* tmm_context_device_c(tmm_context* ctx);
TMM_API size_t tmm_context_device_c_size(tmm_context* ctx); /* elements */
/* mm_handle::set_full_sizes(m, n, k)  — mm_handle.cpp:73-80: make the context's device C hold m x n elements now (grow-only,
 * 1.2x slack, contents discarded on growth like device_vector::resize, device_vector.hpp:92-107).  tmm_gemm with
 * copy_c_back == 0 does this itself; the call exists for callers that size the buffer up front. */
TMM_API int tmm_context_reserve_device_c(tmm_context* ctx, int64_t m, int64_t n);
/* gpu_context::get_stream(i) / get_result_stream()  — gpu_context.cpp:20-57.  The context's own streams as cudaStream_t, for
 * callers that order their device work against the library's: kind TMM_STREAM_COMPUTE (index 0 = the high-priority chain,
 * 1.. = column-block streams), TMM_STREAM_H2D, TMM_STREAM_D2H (the reference's "result stream").  NULL if out of range. */
#define TMM_STREAM_COMPUTE 0
#define TMM_STREAM_H2D 1
#define TMM_STREAM_D2H 2
TMM_API void* tmm_context_stream(tmm_context* ctx, int kind, int index);

/* mm_handle::optimal_tile_sizes / get_max_tile_sizes / get_num_streams / set_num_streams /
 * set_tile_sizes / set_streams_and_tiles  — mm_handle.cpp:36-66,112-148. */
TMM_API int tmm_context_optimal_tile_sizes(tmm_context* ctx, int m, int n, int k, int* tile_m, int* tile_n, int* tile_k);
TMM_API int tmm_context_get_max_tile_sizes(tmm_context* ctx, int* tile_m, int* tile_n, int* tile_k);
TMM_API int tmm_context_get_num_streams(tmm_context* ctx);
TMM_API int tmm_context_set_streams_and_tiles(tmm_context* ctx, int n_streams, int tile_m, int tile_n, int tile_k);
TMM_API int tmm_context_dtype(tmm_context* ctx);

/* gpu::malloc_pinned<T>(N, value) minus the fill  — util.hpp:65-72 (cudaHostAlloc, flags 0). The reference
 * has no matching free helper (its apps leak); tmm_free_pinned is cudaFreeHost. */
TMM_API int tmm_malloc_pinned(size_t bytes, void** out);
/* Additive: large pinned allocations (out-of-core matrices of hundreds of GB) an order of magnitude faster than cudaHostAlloc - anonymous
 * memory on 2 MiB pages, first touched by all cores, one cudaHostRegister (portable).  Zero-filled.  Release with tmm_free_pinned ONLY
 * (not cudaFreeHost); gpu::malloc_pinned keeps the reference's cudaHostAlloc semantics (util.hpp:65-72). */
TMM_API int tmm_malloc_pinned_large(size_t bytes, void** out);
TMM_API int tmm_free_pinned(void* p);
/* gpu::malloc_device / copy_to_device / copy_to_host  — util.hpp:57-63,79-88 */
TMM_API int tmm_malloc_device(size_t bytes, void** out);
TMM_API int tmm_free_device(void* p);
TMM_API int tmm_copy_to_device(const void* host_from, void* device_to, size_t bytes);
TMM_API int tmm_copy_to_host(const void* device_from, void* host_to, size_t bytes);

/* The arithmetic boundary on its own: device-resident GEMM, replaces blas_api::{s,d,c,z}gemm
 * (gpu_blas_api.hpp:194-252 as called from tiled_mm.cpp:181-268).  Device pointers; A and B must be
 * 16-byte aligned with ld*sizeof(elem) a multiple of 16 (TMA); C: any ld >= m.  Runs on `stream`
 * (a cudaStream_t, 0 = default stream) and does not synchronize. */
TMM_API int tmm_device_gemm(int dtype, char trans_a, char trans_b, int64_t m, int64_t n, int64_t k, const void* alpha, const void* a_dev, int64_t ld_a,
                    const void* b_dev, int64_t ld_b, const void* beta, void* c_dev, int64_t ld_c, void* stream);

/* Mixed precision (additive: the reference API has no such type; north_star names BF16 as a tcgen05 kernel family).
 * C (float) = alpha * op(A) * op(B) + beta * C with A and B stored as bfloat16 (the upper 16 bits of an IEEE float), device pointers,
 * column-major, any lda / ldb >= stored rows; FP32 accumulation on the tensor cores.  Runs on `stream`, does not synchronize. */
TMM_API int tmm_device_gemm_bf16(char trans_a, char trans_b, int64_t m, int64_t n, int64_t k, float alpha, const void* a_bf16_dev, int64_t ld_a,
                         const void* b_bf16_dev, int64_t ld_b, float beta, float* c_dev, int64_t ld_c, void* stream);

/* Math mode of the float (TMM_F32) GEMM, process-wide; the counterpart of cublasSetMathMode, which the reference never calls
 * (gpu_blas_handle.hpp:11-17 -> cuBLAS default math = FP32-accurate results).  TMM_MATH_FP32 (default) keeps that accuracy on
 * the tcgen05 tensor cores by splitting every operand into two TF32 numbers (3 MMAs per product, FP32 accumulation in TMEM);
 * TMM_MATH_TF32 is the opt-in fast mode (one TF32 MMA, ~1e-3 relative); TMM_MATH_SIMT forces the FFMA kernel.
 * Also settable with the environment variable TMM_F32_MATH = fp32 | tf32 | simt. */
#define TMM_MATH_SIMT 0
#define TMM_MATH_TF32 1
#define TMM_MATH_FP32 3
TMM_API int tmm_set_f32_math(int mode);
TMM_API int tmm_get_f32_math(void);
/* Math mode of the complex<float> (TMM_C32) GEMM, process-wide.  TMM_CMATH_TC (default): the complex product as a real product of
 * twice the size, (2m x n) = (2m x 2k)(2k x n) over the (re, im) float view of the operands, on the FP32-accurate tcgen05 kernel
 * (3xTF32; FP32 accuracy class, integer data stays exact; 1.8x cuBLAS CGEMM on B200).  TMM_CMATH_SIMT: the complex FMA kernel, true FP32
 * arithmetic in every product.  Environment: TMM_C32_MATH = tc | simt. */
#define TMM_CMATH_SIMT 0
#define TMM_CMATH_TC 3
TMM_API int tmm_set_c32_math(int mode);
TMM_API int tmm_get_c32_math(void);

/* ---- Multi-GPU: C tile-blocks over a p_r x p_c grid of the box's GPUs (no counterpart in the reference, which drives one
 * device; north_star: "partitioned across the 8 B200s of one box by assigning C tile-blocks to GPUs").  GPUs of a grid row
 * share the A row-panel, GPUs of a grid column the B column-panel; every GPU uploads a distinct share of each shared panel
 * over its own PCIe link and the shares are all-gathered over NVLink (NCCL, loaded at run time).  k is never split.
 *
 * (1) One process, many GPUs - the drop-in path: after tmm_context_set_devices(ctx, n, ids) every tmm_gemm(ctx, ...) with
 *     copy_c_back != 0 splits C over n child contexts (one host thread each); a call with copy_c_back == 0 runs on the first device
 *     and tmm_context_device_c() follows it there.  ids == NULL means devices 0..n-1.  The environment variable TMM_DEVICES=n does
 *     the same for every context an application creates, without a code change.
 * (2) One process per GPU (torchrun / MPI): each rank creates a context on its device and joins the grid with
 *     tmm_context_attach_grid(); tmm_gemm(ctx, ...) then computes THIS RANK's C block: m, n are the block's extents, a points
 *     at the rank's rows of op(A) (its A row-panel, full k), b at its columns of op(B), c at its block (ld_c = host ld).
 *     All ranks call tmm_gemm collectively with the same trans / k / beta==0-ness / copy_c_back.  The 128-byte NCCL ids
 *     come from tmm_dist_unique_id() on one rank of each grid row / column and are distributed by the caller. */
TMM_API int tmm_grid_shape(int n_gpus, int* grid_rows, int* grid_cols);                 /* 1->1x1, 2->1x2, 4->2x2, 8->2x4 */
TMM_API int tmm_share_range(int64_t extent, int parts, int index, int64_t* lo, int64_t* hi); /* balanced split used for blocks and upload shares */
TMM_API int tmm_dist_unique_id(void* out128);
TMM_API int tmm_context_attach_grid(tmm_context* ctx, int grid_rows, int grid_cols, int my_row, int my_col, const void* row_id128, const void* col_id128);
TMM_API int tmm_context_grid(tmm_context* ctx, int* grid_rows, int* grid_cols, int* my_row, int* my_col);
TMM_API int tmm_context_set_devices(tmm_context* ctx, int n_devices, const int* device_ids);
TMM_API int tmm_context_num_devices(tmm_context* ctx);
TMM_API tmm_context* tmm_context_child(tmm_context* ctx, int index);
/* cudaMemcpy2DAsync with plain arguments (kind: 1 H2D, 2 D2H, 3 D2D) - what copy_tile_to_device_async / copy_tile_to_host_async
 * (tiled_mm.cpp:45-123) reduce to; lets a binding stage strided panels without a CUDA runtime binding of its own. */
TMM_API int tmm_memcpy_2d_async(void* dst, size_t dpitch_bytes, const void* src, size_t spitch_bytes, size_t width_bytes, size_t height, int kind, void* stream);

/* Box probes (no counterpart in the reference): the roofline denominators of this path measured where the code runs.
 *   tmm_probe_fp64_peak   FP64 tensor (DMMA.8x8x4) issue rate of the current device in TFLOP/s - the ceiling of the DGEMM / ZGEMM kernels
 *   tmm_probe_host_links  pinned H2D and D2H GB/s of each listed device (NULL: 0..n-1) with ALL of them copying both ways at once -
 *                         on a multi-GPU box the links share uplinks / host memory, so the figure per GPU falls with the GPU count */
TMM_API int tmm_probe_fp64_peak(double* tflops);
TMM_API int tmm_probe_host_links(int n_devices, const int* device_ids, size_t bytes_per_direction, double* h2d_gbs, double* d2h_gbs);

/* Introspection for tests and bench.py */
typedef struct tmm_call_stats {
    uint64_t h2d_bytes;      /* bytes moved host->device by the last tmm_gemm */
    uint64_t d2h_bytes;      /* bytes moved device->host by the last tmm_gemm */
    uint64_t kernel_launches;/* kernels launched by the last tmm_gemm */
    uint64_t h2d_copies, d2h_copies;
    double wall_ms;          /* host wall time of the last tmm_gemm */
    double kernel_ms;        /* summed device time of the GEMM kernels of the last call (0 unless profiling is on) */
    int regime;              /* 0 resident (A,B,C fit in HBM), 1 streaming (k-panel ring) */
    int c_blocks, k_chunks;
    uint64_t peer_bytes;     /* bytes this GPU received from peer GPUs over NVLink (GPU grid only) */
} tmm_call_stats;
TMM_API int tmm_context_last_stats(tmm_context* ctx, tmm_call_stats* out);
TMM_API int tmm_context_set_profiling(tmm_context* ctx, int on);          /* time every kernel with CUDA events */
TMM_API int tmm_context_set_device_budget(tmm_context* ctx, size_t bytes); /* cap device memory use (0 = auto); lets tests force the streaming regime */
/* Pure host logic, no GPU needed: the reference's per-dimension tile heuristic (mm_handle.cpp:89-110) and this
 * library's schedule for a call, as a JSON document (regime, phase-1 block and k-chunks, phase-2 blocks / ring geometry). */
TMM_API int tmm_optimal_tile_size(int dim, int max_tile);
TMM_API int tmm_plan_describe(int dtype, char trans_a, char trans_b, int64_t m, int64_t n, int64_t k, int beta_nonzero, int copy_c_back, size_t budget_bytes,
                      int n_streams, int tile_m, int tile_n, int tile_k, int sm_count, char* out, size_t out_size);
TMM_API uint64_t tmm_total_kernel_launches(void);
TMM_API const char* tmm_last_error(void);
TMM_API const char* tmm_version(void);
TMM_API int tmm_device_count(void);

#ifdef __cplusplus
}
#endif
#endif
