// common.cuh — shared device/host helpers for the reductive_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <atomic>

#include "../../include/reductive_b200.h"

namespace rb {

// ---------------------------------------------------------------------------------------------
// host-side error plumbing
// ---------------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;
// Stream-ordered scratch allocation from the library's own memory pool (one per device, never trimmed at
// synchronisation points: the default pool hands its memory back to the driver at every sync, which costs a
// ~0.5 ms re-allocation in the first call after each one).  Free with cudaFreeAsync.
cudaError_t pool_malloc(void **p, size_t bytes, cudaStream_t stream);

#define RB_CUDA_TRY(expr)                                                                        \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            ::rb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,    \
                            __LINE__);                                                           \
            return RB_ERR_CUDA;                                                                  \
        }                                                                                        \
    } while (0)

#define RB_TRY(expr)                        \
    do {                                    \
        rb_status _s = (expr);              \
        if (_s != RB_OK) return _s;         \
    } while (0)

// Debugging aid (RB_DEBUG_SYNC=1): synchronise after every launch and say where on stderr -- finds the kernel that
// hangs or faults.
bool debug_sync_enabled();
void debug_sync_report(const char *file, int line);

// Counts the launch and turns a launch-time error into RB_ERR_CUDA.
#define RB_LAUNCH_CHECK()                                                                        \
    do {                                                                                         \
        ::rb::g_launches.fetch_add(1, std::memory_order_relaxed);                                \
        if (::rb::debug_sync_enabled()) ::rb::debug_sync_report(__FILE__, __LINE__);             \
        cudaError_t _e = cudaGetLastError();                                                     \
        if (_e != cudaSuccess) {                                                                 \
            ::rb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),          \
                            __FILE__, __LINE__);                                                 \
            return RB_ERR_CUDA;                                                                  \
        }                                                                                        \
    } while (0)

static inline size_t ceil_div(size_t a, size_t b) { return (a + b - 1) / b; }

// Number of SMs of the current device (launch geometry is sized from it; 148 on a B200).
static inline int sm_count()
{
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        sms <= 0)
        return 148;
    return sms;
}

// ---------------------------------------------------------------------------------------------
// device arithmetic that must round exactly like the reference's CPU code
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// ndarray numeric_util::unrolled_dot (called at linalg.rs:110-112,136-137,167-168): eight independent
// accumulators with a rounded multiply and a rounded add each, combined pairwise, then the tail.
// __fmul_rn/__fadd_rn are never contracted into FMA by nvcc.
template <typename LoadX, typename LoadY>
__device__ __forceinline__ float unrolled_dot_dev(int n, LoadX lx, LoadY ly)
{
    float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f, p4 = 0.f, p5 = 0.f, p6 = 0.f, p7 = 0.f;
    int i = 0;
    for (; i + 8 <= n; i += 8) {
        p0 = __fadd_rn(p0, __fmul_rn(lx(i + 0), ly(i + 0)));
        p1 = __fadd_rn(p1, __fmul_rn(lx(i + 1), ly(i + 1)));
        p2 = __fadd_rn(p2, __fmul_rn(lx(i + 2), ly(i + 2)));
        p3 = __fadd_rn(p3, __fmul_rn(lx(i + 3), ly(i + 3)));
        p4 = __fadd_rn(p4, __fmul_rn(lx(i + 4), ly(i + 4)));
        p5 = __fadd_rn(p5, __fmul_rn(lx(i + 5), ly(i + 5)));
        p6 = __fadd_rn(p6, __fmul_rn(lx(i + 6), ly(i + 6)));
        p7 = __fadd_rn(p7, __fmul_rn(lx(i + 7), ly(i + 7)));
    }
    float sum = 0.f;
    sum = __fadd_rn(sum, __fadd_rn(p0, p4));
    sum = __fadd_rn(sum, __fadd_rn(p1, p5));
    sum = __fadd_rn(sum, __fadd_rn(p2, p6));
    sum = __fadd_rn(sum, __fadd_rn(p3, p7));
    for (; i < n; i++) sum = __fadd_rn(sum, __fmul_rn(lx(i), ly(i)));
    return sum;
}

// compile-time-length variant over a register array
template <int N>
__device__ __forceinline__ float unrolled_sqnorm_reg(const float (&v)[N])
{
    float p[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    constexpr int FULL = (N / 8) * 8;
#pragma unroll
    for (int i = 0; i < FULL; i++) p[i & 7] = __fadd_rn(p[i & 7], __fmul_rn(v[i], v[i]));
    float sum = 0.f;
    sum = __fadd_rn(sum, __fadd_rn(p[0], p[4]));
    sum = __fadd_rn(sum, __fadd_rn(p[1], p[5]));
    sum = __fadd_rn(sum, __fadd_rn(p[2], p[6]));
    sum = __fadd_rn(sum, __fadd_rn(p[3], p[7]));
#pragma unroll
    for (int i = FULL; i < N; i++) sum = __fadd_rn(sum, __fmul_rn(v[i], v[i]));
    return sum;
}

// ndarray non-contiguous 1-D dot: sum = sum + a*b, sequential.
template <int N>
__device__ __forceinline__ float sequential_sqnorm_reg(const float (&v)[N])
{
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < N; i++) sum = __fadd_rn(sum, __fmul_rn(v[i], v[i]));
    return sum;
}

// linalg.rs:173-174: self_sqn[i] + other_sqn[j] - (dp + dp), left to right in f32.
__device__ __forceinline__ float ref_distance(float xs, float cs, float dp)
{
    return __fsub_rn(__fadd_rn(xs, cs), __fadd_rn(dp, dp));
}

// OrderedFloat "a < b" (ordered-float 2): NaN is the greatest value and equal to itself.
__device__ __forceinline__ bool of_less(float a, float b)
{
    return (a < b) || ((b != b) && (a == a));
}

// `assignment.as_()` truncating store of a code of `width` bytes (primitives.rs:100).
__device__ __forceinline__ void store_code(void *base, int width, long long idx, unsigned v)
{
    switch (width) {
    case 1: reinterpret_cast<uint8_t *>(base)[idx] = (uint8_t)v; break;
    case 2: reinterpret_cast<uint16_t *>(base)[idx] = (uint16_t)v; break;
    case 4: reinterpret_cast<uint32_t *>(base)[idx] = v; break;
    default: reinterpret_cast<uint64_t *>(base)[idx] = (uint64_t)v; break;
    }
}

__device__ __forceinline__ unsigned long long load_code(const void *base, int width, long long idx)
{
    switch (width) {
    case 1: return reinterpret_cast<const uint8_t *>(base)[idx];
    case 2: return reinterpret_cast<const uint16_t *>(base)[idx];
    case 4: return reinterpret_cast<const uint32_t *>(base)[idx];
    default: return reinterpret_cast<const uint64_t *>(base)[idx];
    }
}

#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------------
// kernel launchers (defined in the .cu files; all asynchronous on `stream`)
// ---------------------------------------------------------------------------------------------

// Codebook as the kernels want it, resident on the device.
struct DeviceCodebook {
    const float *quantizers;  // [M,k,dsub]
    const float *cs;          // [M,k]  ||c||^2 in unrolled_dot order (linalg.rs:168)
    size_t M, k, dsub;
};

// encode_exact.cu — exact FP32 SIMT encode.  x: [n, d] with row stride ldx (col stride 1).
// seq_norm != 0 selects the sequential-dot row norm the reference takes for non-contiguous views.
rb_status launch_encode_exact(const DeviceCodebook &cb, const float *x, size_t n, ptrdiff_t ldx,
                              void *codes, int code_width, ptrdiff_t crs, ptrdiff_t ccs, int seq_norm,
                              cudaStream_t stream);
// Same kernel, but it returns immediately unless *gate > gate_thr (gate == nullptr: always runs).  Used by the
// tensor path as its on-device fallback when the list of undecided pairs overflowed.
rb_status launch_encode_exact_gated(const DeviceCodebook &cb, const float *x, size_t n, ptrdiff_t ldx, void *codes,
                                    int code_width, ptrdiff_t crs, ptrdiff_t ccs, int seq_norm, const uint32_t *gate,
                                    uint32_t gate_thr, cudaStream_t stream);
// ||c||^2 for every centroid.
rb_status launch_centroid_norms(const float *quantizers, size_t rows, size_t dsub, float *cs,
                                cudaStream_t stream);
// Exact recheck of flagged (row, subquantizer) pairs (used by the tensor path).
rb_status launch_encode_recheck(const DeviceCodebook &cb, const float *x, ptrdiff_t ldx,
                                const uint32_t *pairs, const uint32_t *n_pairs, uint32_t max_pairs,
                                void *codes, int code_width, ptrdiff_t crs, ptrdiff_t ccs,
                                cudaStream_t stream);

// Pairs the tensor path could not decide: cands holds 4 words per pair (row, subquantizer, chain flags, block flags)
// in `regions` regions of `region_cap` entries, region r holding region_counts[r] of them.  The candidates are the
// centroids 16 b + a with chain a flagged (bit a / 2 + 16 (a % 2) of the chain word) and block b flagged (bit b or
// bit 16 + b of the block word); both words all ones: every centroid.  Decides among them with the reference's
// exact expression tree (first index on ties).
rb_status launch_encode_candidates(const DeviceCodebook &cb, const float *x, ptrdiff_t ldx, const uint32_t *cands,
                                   const uint32_t *region_counts, uint32_t regions, uint32_t region_cap, void *codes,
                                   int code_width, ptrdiff_t crs, ptrdiff_t ccs, cudaStream_t stream);

// The same list for a ROTATED batch (x ~ x0 . R, every component of row i off by at most rowerr[i] + err_floor / *sx):
// decides a pair on the approximate subvector when its candidates are further apart than the rotation error can move
// them, else appends the row to the per-subquantizer buckets (bucket_counts[M], bucket_rows[M][n_cap]) for the exact
// re-rotation below.
rb_status launch_rotated_candidates(const DeviceCodebook &cb, const float *x, ptrdiff_t ldx, const uint32_t *cands,
                                    const uint32_t *region_counts, uint32_t regions, uint32_t region_cap,
                                    const float *rowerr, const float *sx_dev, float err_floor, uint32_t *bucket_counts,
                                    uint32_t *bucket_rows, size_t n_cap, void *codes, int code_width, ptrdiff_t crs,
                                    ptrdiff_t ccs, cudaStream_t stream);

// Flagged rows of a ROTATED batch, bucketed per subquantizer (counts[M], rows[M][n_cap]): re-rotate the subvector
// exactly (reference sgemm order, x0 . r) and re-decide it with the reference's expression tree.
bool rotated_recheck_supported(const DeviceCodebook &cb, size_t d);
rb_status launch_rotated_recheck(const DeviceCodebook &cb, const uint32_t *counts, const uint32_t *rows, size_t n_cap,
                                 const float *x0, ptrdiff_t ldx0, const float *r, size_t d, void *codes, int code_width,
                                 ptrdiff_t crs, ptrdiff_t ccs, cudaStream_t stream);

// gather.cu — reconstruct_batch.  err_flag: device int set to 1 on an out-of-range code.
rb_status launch_gather(const DeviceCodebook &cb, const void *codes, int code_width, size_t n,
                        ptrdiff_t crs, ptrdiff_t ccs, float *out, ptrdiff_t ldo, int *err_flag,
                        cudaStream_t stream);

// pack.cu — strided <-> contiguous helpers
rb_status launch_pack_rows(const float *src, size_t n, size_t d, ptrdiff_t rs, ptrdiff_t cs, float *dst,
                           cudaStream_t stream);
rb_status launch_unpack_rows(const float *src, size_t n, size_t d, float *dst, ptrdiff_t rs, ptrdiff_t cs,
                             cudaStream_t stream);

// project.cu — out = x . R or x . R^T in the reference's accumulation order.
rb_status launch_project(const float *x, size_t n, size_t d, ptrdiff_t rsx, ptrdiff_t csx, const float *r,
                         int transpose_r, float *out, cudaStream_t stream);

// kmeans.cu
// codes: column-major [M][code_pitch] (code_pitch >= n, multiple of 16).  init (or nullptr): packed sums of the rows
// that precede these in the reference's row order (ordered update only) — the chains continue from them.
rb_status launch_kmeans_accumulate(const float *x, size_t n, ptrdiff_t ldx, const uint8_t *codes8,
                                   const uint32_t *codes32, size_t code_pitch, size_t M, size_t k, size_t dsub,
                                   const float *init, float *packed, int ordered, cudaStream_t stream);
// sumsq64 (or nullptr): sum ||x_m||^2 per subquantizer in FP64 from launch_sumsq64, used for the loss instead of the
// packed buffer's float slot.
rb_status launch_kmeans_finalize(const float *packed, size_t M, size_t k, size_t dsub, uint64_t n_total,
                                 float *centroids, float *loss, cudaStream_t stream, const double *sumsq64 = nullptr);
// out[m] = sum over the n rows of ||x[:, m*dsub .. (m+1)*dsub)||^2 in FP64, fixed summation order (deterministic).
rb_status launch_sumsq64(const float *x, size_t n, ptrdiff_t ldx, size_t M, size_t dsub, double *out, cudaStream_t stream);

// opq.cu — OPQ training helpers.  means[d] = column means of x (FP64 accumulation, fixed order);
// out[da, db] = sum_r (a[r, :] - a_sub)^T ((b[r, :] - b_sub) / b_div)  (a_sub / b_sub may be nullptr, b_div 1).
rb_status launch_column_means(const float *x, size_t n, size_t d, ptrdiff_t ldx, float *means, cudaStream_t stream);
rb_status launch_gram(const float *a, ptrdiff_t lda, const float *b, ptrdiff_t ldb, size_t n, size_t da, size_t db,
                      const float *a_sub, const float *b_sub, float b_div, float *out, cudaStream_t stream);

// cabi.cu -- cluster_assignments of all subquantizers into column-major codes with column stride col_stride (elements)
rb_status kmeans_assign_strided(const float *x, size_t n_local, ptrdiff_t ldx, const float *centroids, size_t M, size_t k,
                                size_t dsub, void *codes, ptrdiff_t col_stride, cudaStream_t stream);

// qstore.cu -- caller-side quantized storage: row lookups and the fused decode + dot scan.  codes: dense [n][M] u8,
// allocation padded to a multiple of 16 bytes.  lut: workspace of qstore_lut_floats(M, k, nq) floats.
rb_status launch_qstore_select(const uint8_t *codes, size_t n, size_t M, const unsigned long long *idx, size_t n_idx,
                               uint8_t *out, const float *norms, float *norms_out, int *err, cudaStream_t stream);
rb_status launch_qstore_scale_rows(float *out, ptrdiff_t ors, ptrdiff_t ocs, size_t n, size_t d, const float *norms_sel,
                                   cudaStream_t stream);
rb_status launch_qstore_check_codes(const uint8_t *codes, size_t bytes, size_t k, int *err, cudaStream_t stream);
int qstore_queries_per_pass(size_t M, size_t k, size_t nq);
size_t qstore_lut_floats(size_t M, size_t k, size_t nq);
rb_status launch_qstore_dot(const uint8_t *codes, size_t n, const float *cent, size_t M, size_t k, size_t dsub,
                            const float *qrot, size_t nq, const float *norms, float *lut, float *out, ptrdiff_t out_ld,
                            cudaStream_t stream);

// gram_tc.cu -- the same Gram matrix on tcgen05 (two-limb BF16 operands, FP32 accumulation in segments of 2048 rows)
bool gram_tensor_supported(const float *a, ptrdiff_t lda, const float *b, ptrdiff_t ldb, size_t n, size_t da, size_t db);
rb_status launch_gram_tensor(const float *a, ptrdiff_t lda, const float *b, ptrdiff_t ldb, size_t n, size_t da, size_t db,
                             const float *a_sub, const float *b_sub, float b_div, float *out, cudaStream_t stream);
void set_gram_algo(int algo);

// Streaming ordered update for training loops (kmeans.cu): x laid out once as subquantizer-major slabs
// (slab_floats(n, d) floats), then per iteration one pass that streams the slabs; sums bit-identical to the chain path.
bool stream_update_supported(size_t k, size_t dsub);
bool kmeans_stream_enabled();  // rb_set_kmeans_update(3): training loops use the streaming update (experimental, slower)
int kmeans_update_mode();      // 0: shared-memory atomics, 1 / 2: ordered (sort + chains), 3: ordered, streaming slabs
size_t slab_floats(size_t n, size_t d);
rb_status launch_build_slabs(const float *x, size_t n, ptrdiff_t ldx, size_t M, size_t dsub, float *slabs, cudaStream_t stream);
rb_status launch_ordered_stream(const float *slabs, size_t n, const uint8_t *codes, size_t code_pitch, size_t M, size_t k,
                                size_t dsub, float *packed, cudaStream_t stream);

// vector_ops.cu — single-vector paths (latency only).
rb_status launch_quantize_vector(const DeviceCodebook &cb, const float *projection, const float *x,
                                 ptrdiff_t sx, void *codes, int code_width, ptrdiff_t cstride,
                                 float *scratch, cudaStream_t stream);
rb_status launch_reconstruct_vector(const DeviceCodebook &cb, const float *projection, const void *codes,
                                    int code_width, ptrdiff_t cstride, float *out, ptrdiff_t ostride,
                                    float *scratch, int *err_flag, cudaStream_t stream);

}  // namespace rb
