// opq.cu — device pieces of Opq / GaussianOpq TRAINING (SURVEY 8f "next" rows f1 / f3).
//
// Replaces  Covariance::covariance     src/linalg.rs:23-44    (mean over rows, centred Gram / (n - 1))
//           the Gram matrix  X^T . Y^  src/pq/opq.rs:187      (input of the Procrustes SVD)
//           Opq::train_iteration       src/pq/opq.rs:161-189  (rb_opq_train_iteration in cabi.cu drives these kernels
//                                                              together with the projection, k-means and encode ones)
// The d x d eigendecomposition / SVD stay on host LAPACK as in the reference (opq.rs:123,187; north_star).
// Large Gram matrices (n >= 4096, 16 <= d <= 512) run on the tensor cores (gram_tc.cu); the kernel below serves the rest.
//
// gram_kernel: out[i, j] = sum_r (a[r, i] - a_sub[i]) * ((b[r, j] - b_sub[j]) / b_div).  Reduction over the n rows is
// split over the grid's z dimension into per-split partial matrices that a second kernel adds IN ORDER (deterministic;
// no float atomics).  FP32 FMA on CUDA cores: 2 n d^2 flop, not on the hot path (training only); a 64 x 64 output
// tile per block, 4 x 4 outputs per thread, both operand tiles staged through shared memory with coalesced row reads.
#include <atomic>

#include "common.cuh"

namespace rb {
namespace {

constexpr int kGT = 64;    // output tile edge
constexpr int kGR = 32;    // rows per shared-memory step

__global__ void __launch_bounds__(256)
column_sum_partial_kernel(const float *__restrict__ x, long long n, long long ldx, int d, long long rows_per_split,
                          double *__restrict__ partial)
{
    // block (column tile of 256 columns, row split): fixed row range, FP64 running sum per column, fixed order; eight
    // independent (coalesced) row loads in flight per thread
    const int col = blockIdx.x * 256 + threadIdx.x;
    const long long r0 = (long long)blockIdx.y * rows_per_split, r1 = min(n, r0 + rows_per_split);
    if (col >= d) return;
    double acc = 0.0;
    long long r = r0;
    for (; r + 8 <= r1; r += 8) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; e++) v[e] = __ldg(x + (r + e) * ldx + col);
#pragma unroll
        for (int e = 0; e < 8; e++) acc += (double)v[e];
    }
    for (; r < r1; r++) acc += (double)__ldg(x + r * ldx + col);
    partial[(size_t)blockIdx.y * d + col] = acc;
}

__global__ void __launch_bounds__(256)
column_mean_final_kernel(const double *__restrict__ partial, int d, int n_splits, double inv_n, float *__restrict__ means)
{
    // one warp per column: lane l adds the splits l, l + 32, ... in order, then a fixed shuffle tree (deterministic)
    const int col = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (col >= d) return;
    double s = 0.0;
    for (int p = lane; p < n_splits; p += 32) s += partial[(size_t)p * d + col];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) means[col] = (float)(s * inv_n);
}

__global__ void __launch_bounds__(256)
gram_partial_kernel(const float *__restrict__ a, long long lda, const float *__restrict__ b, long long ldb, long long n, int da,
                    int db, const float *__restrict__ a_sub, const float *__restrict__ b_sub, float b_div, int n_splits,
                    float *__restrict__ partial)
{
    __shared__ __align__(16) float As[kGR][kGT + 4], Bs[kGR][kGT + 4];
    const int i0 = blockIdx.y * kGT, j0 = blockIdx.x * kGT;
    const long long per = ((n + n_splits - 1) / n_splits + kGR - 1) / kGR * kGR;
    const long long r0 = (long long)blockIdx.z * per, r1 = min(n, r0 + per);
    const int ti = threadIdx.x / 16, tj = threadIdx.x % 16;  // 16 x 16 threads, 4 x 4 outputs each
    float acc[4][4] = {};
    for (long long rb = r0; rb < r1; rb += kGR) {
        // stage kGR rows x 64 columns of both operands (centred / scaled on the way in, like the reference's
        // `centered` and `centered.map(|v| v / normalization)`)
        for (int e = threadIdx.x; e < kGR * kGT; e += 256) {
            const int r = e / kGT, c = e % kGT;
            const long long row = rb + r;
            float va = 0.f, vb = 0.f;
            if (row < r1) {
                if (i0 + c < da) va = __fsub_rn(__ldg(a + row * lda + i0 + c), a_sub ? a_sub[i0 + c] : 0.f);
                if (j0 + c < db) {
                    vb = __fsub_rn(__ldg(b + row * ldb + j0 + c), b_sub ? b_sub[j0 + c] : 0.f);
                    if (b_div != 1.f) vb = __fdiv_rn(vb, b_div);
                }
            }
            As[r][c] = va;
            Bs[r][c] = vb;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < kGR; r++) {
            const float4 av = *reinterpret_cast<const float4 *>(&As[r][4 * ti]);
            const float4 bv = *reinterpret_cast<const float4 *>(&Bs[r][4 * tj]);
            const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int p = 0; p < 4; p++)
#pragma unroll
                for (int q = 0; q < 4; q++) acc[p][q] = fmaf(ar[p], br[q], acc[p][q]);
        }
        __syncthreads();
    }
    float *out = partial + (size_t)blockIdx.z * da * db;
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int i = i0 + 4 * ti + p, j = j0 + 4 * tj + q;
            if (i < da && j < db) out[(size_t)i * db + j] = acc[p][q];
        }
}

__global__ void gram_final_kernel(const float *__restrict__ partial, size_t len, int n_splits, float *__restrict__ out)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= len) return;
    float s = 0.f;
    for (int p = 0; p < n_splits; p++) s += partial[(size_t)p * len + e];  // fixed order
    out[e] = s;
}

}  // namespace

rb_status launch_column_means(const float *x, size_t n, size_t d, ptrdiff_t ldx, float *means, cudaStream_t stream)
{
    if (n == 0 || d == 0) return RB_OK;
    // row splits of at least 512 rows, enough of them to fill the GPU (the sums of a split and the order in which the
    // splits are added are fixed: deterministic)
    size_t rows_per_split = 512;
    if (ceil_div(n, rows_per_split) > 4096) rows_per_split = ceil_div(n, (size_t)4096);
    const int splits = (int)ceil_div(n, rows_per_split);
    double *partial = nullptr;
    RB_CUDA_TRY(pool_malloc((void **)&partial, (size_t)splits * d * sizeof(double), stream));
    column_sum_partial_kernel<<<dim3((unsigned)ceil_div(d, (size_t)256), (unsigned)splits), 256, 0, stream>>>(
        x, (long long)n, (long long)ldx, (int)d, (long long)rows_per_split, partial);
    column_mean_final_kernel<<<(unsigned)ceil_div(d, (size_t)8), 256, 0, stream>>>(partial, (int)d, splits, 1.0 / (double)n, means);
    cudaFreeAsync(partial, stream);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

static std::atomic<int> g_gram_algo{0};  // rb_set_gram_algo: 0 auto, 1 FP32 CUDA cores, 2 tensor cores (gram_tc.cu)
void set_gram_algo(int algo) { g_gram_algo.store(algo); }

rb_status launch_gram(const float *a, ptrdiff_t lda, const float *b, ptrdiff_t ldb, size_t n, size_t da, size_t db,
                      const float *a_sub, const float *b_sub, float b_div, float *out, cudaStream_t stream)
{
    if (da == 0 || db == 0) return RB_OK;
    const int algo = g_gram_algo.load();
    const bool tc_ok = gram_tensor_supported(a, lda, b, ldb, n, da, db);
    if (algo == 2 && !tc_ok) {
        set_error("tensor-core Gram does not cover this call (n=%zu, %zu x %zu)", n, da, db);
        return RB_ERR_UNSUPPORTED;
    }
    if (algo != 1 && tc_ok) return launch_gram_tensor(a, lda, b, ldb, n, da, db, a_sub, b_sub, b_div, out, stream);
    const size_t tiles = ceil_div(da, (size_t)kGT) * ceil_div(db, (size_t)kGT);
    // enough row splits to fill the GPU a few times over, at least 4096 rows each
    size_t splits = ceil_div((size_t)sm_count() * 4, tiles);
    if (splits > ceil_div(n, (size_t)4096)) splits = ceil_div(n, (size_t)4096);
    if (splits < 1) splits = 1;
    float *partial = nullptr;
    RB_CUDA_TRY(pool_malloc((void **)&partial, splits * da * db * sizeof(float), stream));
    gram_partial_kernel<<<dim3((unsigned)ceil_div(db, (size_t)kGT), (unsigned)ceil_div(da, (size_t)kGT), (unsigned)splits), 256, 0,
                          stream>>>(a, (long long)lda, b, (long long)ldb, (long long)n, (int)da, (int)db, a_sub, b_sub, b_div,
                                    (int)splits, partial);
    gram_final_kernel<<<(unsigned)ceil_div(da * db, (size_t)256), 256, 0, stream>>>(partial, da * db, (int)splits, out);
    cudaFreeAsync(partial, stream);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

}  // namespace rb
