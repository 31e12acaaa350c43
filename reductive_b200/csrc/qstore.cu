// Caller-side quantized storage: the consumer of a trained Pq<f32> in finalfusion is a
// "quantized array" = quantizer + [n, M] u8 codes + optional per-row norms, whose lookups are
//     embedding(i) = reconstruct(codes[i]) * norm[i]
// built on Reconstruct (reference: src/pq/traits.rs:102-156, src/pq/pq.rs:303-347; the storage type itself lives in
// the finalfusion crate, outside /root/reference).  Kernels here:
//   qstore_select_kernel      codes / norms of the requested rows -> a dense batch (then gather.cu decodes it)
//   qstore_scale_rows_kernel  reconstruction * norm, one rounded multiply per element like ndarray's `*=`
//   qstore_lut_kernel         per query and subquantizer: dot products of the query's subvector with the 2^bits centroids
//   qstore_scan_kernel        fused decode + dot: score(q, i) = norm[i] * sum_m lut[q][m][codes[i][m]] -- the [n, d]
//                             reconstruction is never materialised; HBM traffic = the codes, once per batch of QB queries
#include "common.cuh"

namespace rb {
namespace {

__global__ void __launch_bounds__(256)
qstore_select_kernel(const uint8_t *__restrict__ codes, unsigned long long n, int M, const unsigned long long *__restrict__ idx,
                     unsigned long long n_idx, uint8_t *__restrict__ out, const float *__restrict__ norms,
                     float *__restrict__ norms_out, int *__restrict__ err)
{
    const unsigned long long total = n_idx * (unsigned long long)M;
    for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long i = e / (unsigned)M;
        const int m = (int)(e % (unsigned)M);
        const unsigned long long row = idx[i];
        if (row >= n) {  // the reference would panic on the out-of-bounds row index
            *err = 1;
            out[e] = 0;
            continue;
        }
        out[e] = codes[row * (unsigned)M + m];
        if (m == 0 && norms_out) norms_out[i] = norms[row];
    }
}

__global__ void __launch_bounds__(256)
qstore_scale_rows_kernel(float *__restrict__ out, long long ors, long long ocs, unsigned long long n, int d,
                         const float *__restrict__ norms_sel)
{
    const unsigned long long total = n * (unsigned long long)d;
    for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long i = e / (unsigned)d;
        const int c = (int)(e % (unsigned)d);
        float *p = out + (long long)i * ors + (long long)c * ocs;
        *p = __fmul_rn(*p, norms_sel[i]);
    }
}

__global__ void __launch_bounds__(256)
qstore_max_code_kernel(const uint8_t *__restrict__ codes, unsigned long long bytes, unsigned k, int *__restrict__ err)
{
    bool bad = false;
    for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < bytes;
         e += (unsigned long long)gridDim.x * blockDim.x)
        bad |= codes[e] >= k;
    if (__syncthreads_or(bad) && threadIdx.x == 0) *err = 1;
}

// lut layout: [query batch][m][j][QB] so that one vector load serves QB queries.  Sequential fused multiply-adds over
// the subvector, t = 0 .. dsub - 1.
__global__ void __launch_bounds__(256)
qstore_lut_kernel(const float *__restrict__ qrot, int nq, int d, const float *__restrict__ cent, int M, int k, int dsub, int QB,
                  float *__restrict__ lut)
{
    const int n_qb = (nq + QB - 1) / QB;
    const long long total = (long long)n_qb * M * k * QB;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(e % QB);
        const long long r = e / QB;
        const int j = (int)(r % k);
        const int m = (int)((r / k) % M);
        const int qb = (int)(r / ((long long)k * M));
        const int qi = qb * QB + q;
        float acc = 0.f;
        if (qi < nq) {
            const float *qv = qrot + (long long)qi * d + (long long)m * dsub;
            const float *cv = cent + ((long long)m * k + j) * dsub;
            for (int t = 0; t < dsub; t++) acc = fmaf(qv[t], cv[t], acc);
        }
        lut[e] = acc;
    }
}

constexpr int kScanThreads = 512;  // one row per thread and tile

template <int QB>
__device__ __forceinline__ void lut_add(float (&acc)[QB], const float *base, unsigned entry)
{
    if constexpr (QB == 1) {
        acc[0] = __fadd_rn(acc[0], base[entry]);
    } else if constexpr (QB == 2) {
        const float2 v = reinterpret_cast<const float2 *>(base)[entry];
        acc[0] = __fadd_rn(acc[0], v.x);
        acc[1] = __fadd_rn(acc[1], v.y);
    } else {
#pragma unroll
        for (int h = 0; h < QB / 4; h++) {
            const float4 v = reinterpret_cast<const float4 *>(base)[entry * (QB / 4) + h];
            acc[4 * h] = __fadd_rn(acc[4 * h], v.x);
            acc[4 * h + 1] = __fadd_rn(acc[4 * h + 1], v.y);
            acc[4 * h + 2] = __fadd_rn(acc[4 * h + 2], v.z);
            acc[4 * h + 3] = __fadd_rn(acc[4 * h + 3], v.w);
        }
    }
}

// grid = (row-tile CTAs, query batches).  Shared memory: the batch's lookup table (SMEM_LUT) and two code tiles filled
// by 16-byte asynchronous copies one tile ahead.  `codes` is padded to a multiple of 16 bytes by the store.
template <int QB, bool SMEM_LUT>
__global__ void __launch_bounds__(kScanThreads, 1)
qstore_scan_kernel(const uint8_t *__restrict__ codes, unsigned long long n, int M, int k, const float *__restrict__ lut,
                   int nq, const float *__restrict__ norms, float *__restrict__ out, long long out_ld)
{
    extern __shared__ __align__(16) unsigned char sm[];
    const size_t lut_floats = (size_t)M * k * QB;
    const float *lut_b = lut + (size_t)blockIdx.y * lut_floats;
    float *slut = reinterpret_cast<float *>(sm);
    unsigned char *stile = sm + (SMEM_LUT ? ((lut_floats * sizeof(float) + 15) & ~(size_t)15) : 0);
    const unsigned tile_bytes = (unsigned)kScanThreads * (unsigned)M;  // multiple of 16
    const unsigned long long total_bytes = (n * (unsigned long long)M + 15ull) & ~15ull;
    const unsigned long long n_tiles = (n + kScanThreads - 1) / kScanThreads;

    auto fill = [&](unsigned long long tile, int buf) {
        const unsigned long long g0 = tile * tile_bytes;
        for (unsigned o = threadIdx.x * 16u; o < tile_bytes; o += kScanThreads * 16u) {
            if (g0 + o < total_bytes) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(stile + (size_t)buf * tile_bytes + o);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(codes + g0 + o) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    unsigned long long tile = blockIdx.x;
    if (tile < n_tiles) fill(tile, 0);
    if constexpr (SMEM_LUT) {
        if ((lut_floats & 3) == 0) {
            for (size_t i = threadIdx.x * 4; i < lut_floats; i += kScanThreads * 4)
                *reinterpret_cast<float4 *>(slut + i) = __ldg(reinterpret_cast<const float4 *>(lut_b + i));
        } else {  // k = 2 with an odd number of subquantizers
            for (size_t i = threadIdx.x; i < lut_floats; i += kScanThreads) slut[i] = __ldg(lut_b + i);
        }
    }
    const float *table = SMEM_LUT ? slut : lut_b;
    const int q0 = blockIdx.y * QB;
    int buf = 0;
    for (; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
        const unsigned long long next = tile + gridDim.x;
        if (next < n_tiles) fill(next, buf ^ 1);
        else asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();  // the tile (and, the first time, the table) is visible to every thread
        const unsigned long long row = tile * kScanThreads + threadIdx.x;
        if (row < n) {
            const unsigned char *rc = stile + (size_t)buf * tile_bytes + (size_t)threadIdx.x * M;
            float acc[QB];
#pragma unroll
            for (int q = 0; q < QB; q++) acc[q] = 0.f;
            for (int m = 0; m < M; m++) lut_add<QB>(acc, table, (unsigned)m * (unsigned)k + rc[m]);  // m ascending
            const float nv = norms ? norms[row] : 1.f;
#pragma unroll
            for (int q = 0; q < QB; q++)
                if (q0 + q < nq) out[(long long)(q0 + q) * out_ld + (long long)row] = norms ? __fmul_rn(acc[q], nv) : acc[q];
        }
        __syncthreads();  // everyone is done with this buffer before it is refilled two iterations later
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

}  // namespace

rb_status launch_qstore_select(const uint8_t *codes, size_t n, size_t M, const unsigned long long *idx, size_t n_idx,
                               uint8_t *out, const float *norms, float *norms_out, int *err, cudaStream_t stream)
{
    if (n_idx == 0) return RB_OK;
    const size_t total = n_idx * M;
    const unsigned blocks = (unsigned)std::min<size_t>(ceil_div(total, (size_t)256), (size_t)sm_count() * 8);
    qstore_select_kernel<<<blocks, 256, 0, stream>>>(codes, n, (int)M, idx, n_idx, out, norms, norms_out, err);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

rb_status launch_qstore_scale_rows(float *out, ptrdiff_t ors, ptrdiff_t ocs, size_t n, size_t d, const float *norms_sel,
                                   cudaStream_t stream)
{
    if (n == 0 || d == 0) return RB_OK;
    const unsigned blocks = (unsigned)std::min<size_t>(ceil_div(n * d, (size_t)256), (size_t)sm_count() * 8);
    qstore_scale_rows_kernel<<<blocks, 256, 0, stream>>>(out, (long long)ors, (long long)ocs, n, (int)d, norms_sel);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

rb_status launch_qstore_check_codes(const uint8_t *codes, size_t bytes, size_t k, int *err, cudaStream_t stream)
{
    if (bytes == 0 || k >= 256) return RB_OK;
    const unsigned blocks = (unsigned)std::min<size_t>(ceil_div(bytes, (size_t)256), (size_t)sm_count() * 8);
    qstore_max_code_kernel<<<blocks, 256, 0, stream>>>(codes, bytes, (unsigned)k, err);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

// queries per pass: the largest of 8 / 4 / 2 / 1 whose table fits shared memory beside the two code tiles (and that the
// call can fill); 0: not even one query's table fits -> the table stays in global memory (L1 / L2 serve it)
int qstore_queries_per_pass(size_t M, size_t k, size_t nq)
{
    const size_t budget = 200 * 1024 - 2 * (size_t)kScanThreads * M;
    for (int qb = 8; qb >= 1; qb >>= 1) {
        if ((size_t)qb > nq && qb > 1) continue;
        if (M * k * qb * sizeof(float) + 16 <= budget && (size_t)kScanThreads * M * 2 < 200 * 1024) return qb;
    }
    return 0;
}

size_t qstore_lut_floats(size_t M, size_t k, size_t nq)
{
    int qb = qstore_queries_per_pass(M, k, nq);
    if (qb == 0) qb = 1;
    return ceil_div(nq, (size_t)qb) * M * k * qb;
}

rb_status launch_qstore_dot(const uint8_t *codes, size_t n, const float *cent, size_t M, size_t k, size_t dsub,
                            const float *qrot, size_t nq, const float *norms, float *lut, float *out, ptrdiff_t out_ld,
                            cudaStream_t stream)
{
    if (n == 0 || nq == 0) return RB_OK;
    if ((size_t)kScanThreads * M * 2 >= 200 * 1024) {
        set_error("decode + dot scan: %zu subquantizers do not fit the code tiles in shared memory", M);
        return RB_ERR_UNSUPPORTED;
    }
    int qb = qstore_queries_per_pass(M, k, nq);
    const bool smem_lut = qb != 0;
    if (!smem_lut) qb = 1;
    const size_t n_qb = ceil_div(nq, (size_t)qb);
    {
        const size_t total = n_qb * M * k * qb;
        const unsigned blocks = (unsigned)std::min<size_t>(ceil_div(total, (size_t)256), (size_t)sm_count() * 16);
        qstore_lut_kernel<<<blocks, 256, 0, stream>>>(qrot, (int)nq, (int)(M * dsub), cent, (int)M, (int)k, (int)dsub, qb, lut);
        RB_LAUNCH_CHECK();
    }
    const size_t tiles = ceil_div(n, (size_t)kScanThreads);
    // all query batches of a wave read the same code tiles: keep the row-tile CTAs of one batch together (x fastest)
    unsigned ctas_x = (unsigned)std::min<size_t>(tiles, (size_t)sm_count());
    if (n_qb < (size_t)sm_count()) ctas_x = (unsigned)std::min<size_t>(tiles, std::max<size_t>(1, (size_t)sm_count() / n_qb));
    const size_t smem = (smem_lut ? ((M * k * qb * sizeof(float) + 15) & ~(size_t)15) : 0) + 2 * (size_t)kScanThreads * M;
    const dim3 grid(ctas_x, (unsigned)n_qb);
#define RB_SCAN(QB, SL)                                                                                              \
    do {                                                                                                             \
        RB_CUDA_TRY(cudaFuncSetAttribute(qstore_scan_kernel<QB, SL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        qstore_scan_kernel<QB, SL><<<grid, kScanThreads, smem, stream>>>(codes, n, (int)M, (int)k, lut, (int)nq, norms, out, \
                                                                         (long long)out_ld);                       \
    } while (0)
    if (!smem_lut) RB_SCAN(1, false);
    else if (qb == 8) RB_SCAN(8, true);
    else if (qb == 4) RB_SCAN(4, true);
    else if (qb == 2) RB_SCAN(2, true);
    else RB_SCAN(1, true);
#undef RB_SCAN
    RB_LAUNCH_CHECK();
    return RB_OK;
}

}  // namespace rb
