// kmeans.cu — update_centroids as a segmented sum + finalize (sm_100a).
//
// Replaces  kmeans::update_centroids    src/kmeans.rs:166-198   (scatter-add, divide, empties stay zero)
//           kmeans::mean_squared_error  src/kmeans.rs:330-360   (from accumulated moments)
// for all M subquantizers of a product quantizer in one launch per phase.
//
// accumulate: reads x [n, d] once (4*d bytes/row) and the codes [n, M] (M bytes/row), writes the packed
//   accumulator  sums [M,k,dsub] | counts [M,k] | sumsq [M]  — which is also the all-reduce payload of
//   data-parallel k-means.  Bound: HBM read bandwidth; the scatter itself goes to per-block shared-memory
//   accumulators (one subquantizer's [k, dsub] slice per block) that are flushed once with global
//   reductions, so contention on global atomics is O(blocks), not O(rows).
// finalize: M*k*dsub elements; divides non-empty clusters with a true IEEE division (`centroid /=
//   count`, kmeans.rs:195), leaves empty clusters at zero (kmeans.rs:181,194), and evaluates
//   sum ||x - c_a||^2 = sum ||x||^2 - 2 sum_j c_j.S_j + sum_j n_j ||c_j||^2 in FP64.
//
// Parity: the reference adds rows sequentially in row order; this adds them in parallel, so sums agree
// to FP32 summation-order error (tests bound trained centroids at 1e-4 relative, as north_star states).
// Counts follow the reference's f32 `+= 1.0` (exact to 2^24 per cluster, kmeans.rs:188).
#include "common.cuh"

namespace rb {

namespace {

constexpr int kAccThreads = 256;

template <bool SMEM_ACC, typename CodeT>
__global__ void __launch_bounds__(kAccThreads)
accumulate_kernel(const float *__restrict__ x, long long n, long long ldx, const CodeT *__restrict__ codes, int M,
                  int k, int dsub, float *__restrict__ packed, long long rows_per_block)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int m = blockIdx.y;
    const int pad = dsub | 1;  // odd row pitch spreads the banks
    float *ssum = reinterpret_cast<float *>(smem_raw);  // [k][pad]
    int *scnt = reinterpret_cast<int *>(ssum + (SMEM_ACC ? (size_t)k * pad : 0));  // [k]

    float *gsum = packed + (size_t)m * k * dsub;
    float *gcnt = packed + (size_t)M * k * dsub + (size_t)m * k;
    float *gsq = packed + (size_t)M * k * dsub + (size_t)M * k + m;

    if constexpr (SMEM_ACC) {
        for (int i = threadIdx.x; i < k * pad; i += kAccThreads) ssum[i] = 0.f;
        for (int i = threadIdx.x; i < k; i += kAccThreads) scnt[i] = 0;
        __syncthreads();
    }

    const long long r0 = (long long)blockIdx.x * rows_per_block;
    const long long r1 = min(n, r0 + rows_per_block);

    // thread -> (row lane rl, component t): consecutive lanes read consecutive floats of a subvector
    const int rows_par = dsub <= kAccThreads ? kAccThreads / dsub : 1;
    const int t_step = dsub <= kAccThreads ? dsub : kAccThreads;
    const int rl = dsub <= kAccThreads ? (int)threadIdx.x / dsub : 0;
    double sq = 0.0;
    if (rl < rows_par) {
        for (int t = dsub <= kAccThreads ? (int)threadIdx.x % dsub : (int)threadIdx.x; t < dsub; t += t_step) {
            for (long long row = r0 + rl; row < r1; row += rows_par) {
                const int code = (int)codes[row * M + m];
                const float v = __ldg(x + row * ldx + (long long)m * dsub + t);
                sq += (double)v * (double)v;
                if constexpr (SMEM_ACC) {
                    atomicAdd(ssum + (size_t)code * pad + t, v);
                    if (t == 0) atomicAdd(scnt + code, 1);
                } else {
                    atomicAdd(gsum + (size_t)code * dsub + t, v);
                    if (t == 0) atomicAdd(gcnt + code, 1.0f);
                }
            }
            if (dsub <= kAccThreads) break;
        }
    }

    // block reduction of the squared-norm partials (FP64), one float atomic per block
    __shared__ double red[kAccThreads / 32];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < kAccThreads / 32; w++) s += red[w];
        atomicAdd(gsq, (float)s);
    }

    if constexpr (SMEM_ACC) {
        // flush: one global reduction per touched accumulator per block
        for (int i = threadIdx.x; i < k * dsub; i += kAccThreads) {
            const int j = i / dsub, t = i - j * dsub;
            const float s = ssum[(size_t)j * pad + t];
            if (s != 0.f) atomicAdd(gsum + i, s);
        }
        for (int j = threadIdx.x; j < k; j += kAccThreads) {
            const int c = scnt[j];
            if (c != 0) atomicAdd(gcnt + j, (float)c);
        }
    }
}

__global__ void __launch_bounds__(256)
finalize_kernel(const float *__restrict__ packed, int M, int k, int dsub, double inv_len, float *__restrict__ centroids,
                float *__restrict__ loss)
{
    const int m = blockIdx.x;
    const float *gsum = packed + (size_t)m * k * dsub;
    const float *gcnt = packed + (size_t)M * k * dsub + (size_t)m * k;
    const float sumsq = packed[(size_t)M * k * dsub + (size_t)M * k + m];
    float *cen = centroids + (size_t)m * k * dsub;

    double acc = 0.0;
    for (int i = threadIdx.x; i < k * dsub; i += blockDim.x) {
        const int j = i / dsub;
        // the reference counts in f32 with `+= 1.0` (kmeans.rs:188), which saturates at 2^24
        const float cnt = fminf(gcnt[j], 16777216.0f);
        const float s = gsum[i];
        float c = 0.f;                       // kmeans.rs:181: empty clusters stay at the zero vector
        if (cnt > 0.f) c = __fdiv_rn(s, cnt);  // kmeans.rs:194-196
        cen[i] = c;
        acc += (double)gcnt[j] * (double)c * (double)c - 2.0 * (double)c * (double)s;
    }
    if (loss == nullptr) return;
    __shared__ double red[8];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = (double)sumsq;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += red[w];
        if (s < 0.0) s = 0.0;
        loss[m] = (float)(s * inv_len);  // kmeans.rs:359: sse / (n * dsub)
    }
}

template <typename CodeT>
rb_status launch_acc_t(const float *x, size_t n, ptrdiff_t ldx, const CodeT *codes, size_t M, size_t k, size_t dsub,
                       float *packed, cudaStream_t stream)
{
    const size_t pad = dsub | 1;
    const size_t smem = k * pad * sizeof(float) + k * sizeof(int);
    const bool smem_acc = smem <= 160 * 1024;
    // ~4 blocks per SM across all subquantizers, at least 1024 rows per block
    size_t chunks = ceil_div((size_t)148 * 4, M);
    size_t rows_per_block = ceil_div(n, chunks);
    if (rows_per_block < 1024) rows_per_block = 1024;
    chunks = ceil_div(n, rows_per_block);
    dim3 grid((unsigned)chunks, (unsigned)M);
    if (smem_acc) {
        auto kern = accumulate_kernel<true, CodeT>;
        if (smem > 48 * 1024)
            RB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kAccThreads, smem, stream>>>(x, (long long)n, (long long)ldx, codes, (int)M, (int)k, (int)dsub,
                                                 packed, (long long)rows_per_block);
    } else {
        accumulate_kernel<false, CodeT><<<grid, kAccThreads, 0, stream>>>(
            x, (long long)n, (long long)ldx, codes, (int)M, (int)k, (int)dsub, packed, (long long)rows_per_block);
    }
    RB_LAUNCH_CHECK();
    return RB_OK;
}

}  // namespace

rb_status launch_kmeans_accumulate(const float *x, size_t n, ptrdiff_t ldx, const uint8_t *codes8,
                                   const uint32_t *codes32, size_t M, size_t k, size_t dsub, float *packed,
                                   cudaStream_t stream)
{
    RB_CUDA_TRY(cudaMemsetAsync(packed, 0, rb_kmeans_packed_len(M, k, dsub) * sizeof(float), stream));
    if (n == 0) return RB_OK;
    if (codes8) return launch_acc_t<uint8_t>(x, n, ldx, codes8, M, k, dsub, packed, stream);
    return launch_acc_t<uint32_t>(x, n, ldx, codes32, M, k, dsub, packed, stream);
}

rb_status launch_kmeans_finalize(const float *packed, size_t M, size_t k, size_t dsub, uint64_t n_total,
                                 float *centroids, float *loss, cudaStream_t stream)
{
    const double len = (double)n_total * (double)dsub;
    finalize_kernel<<<(unsigned)M, 256, 0, stream>>>(packed, (int)M, (int)k, (int)dsub, len > 0 ? 1.0 / len : 0.0,
                                                     centroids, loss);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

}  // namespace rb
