// gather.cu — reconstruct_batch as a coalesced codebook gather (sm_100a).
//
// Replaces  primitives::reconstruct_batch_into  src/pq/primitives.rs:150-173
//        -> primitives::reconstruct_into        src/pq/primitives.rs:110-148
//   out[i, m*dsub .. (m+1)*dsub] = quantizers[m, codes[i, m], ..]
// A pure copy, so bit-exactness is structural.  Bound: HBM write bandwidth — algorithmic bytes per
// vector = M*sizeof(code) read + 4*d written (C2: 1 230 B).
//
// Layout of the work: the output is cut into "pieces" of PW floats (PW = 4 when dsub % 4 == 0, 2 when
// dsub is even, else 1) so that a piece never straddles two subquantizers and every global store is a
// PW*4-byte vector.  Consecutive lanes own consecutive pieces of a row -> each warp store instruction
// writes 32*PW*4 contiguous bytes.  A block owns a contiguous range of subquantizers ("column group")
// whose codebook slice is staged in shared memory once and then reused for a strip of rows, so the
// scattered 4/8/16-byte centroid reads hit shared memory (2-4 wavefronts per warp) instead of L1
// (up to 32 wavefronts); it walks its strip with a grid-stride loop.  When even one subquantizer's
// slice does not fit in shared memory the block reads centroids through the read-only path instead.
#include "common.cuh"

namespace rb {

namespace {

constexpr int kGatherThreads = 512;
constexpr int kUnroll = 8;

template <int PW>
struct VecT;
template <>
struct VecT<1> { using type = float; };
template <>
struct VecT<2> { using type = float2; };
template <>
struct VecT<4> { using type = float4; };

template <int PW>
__device__ __forceinline__ void store_streaming(float *dst, typename VecT<PW>::type v)
{
    if constexpr (PW == 4) {
        __stcs(reinterpret_cast<float4 *>(dst), v);
    } else if constexpr (PW == 2) {
        __stcs(reinterpret_cast<float2 *>(dst), v);
    } else {
        __stcs(dst, v);
    }
}

// grid.x = row strips, grid.y = column groups.
//   piece p (local to the group) -> (m_local = p / ppsq, t = (p % ppsq) * PW), ppsq = dsub / PW.
template <int PW, bool SMEM_CB, int CW>
__global__ void __launch_bounds__(kGatherThreads)
gather_kernel(const float *__restrict__ quantizers, int k, int dsub, int M, const void *__restrict__ codes,
              int code_width, long long n, long long crs, long long ccs, float *__restrict__ out, long long ldo,
              int m_per_group, long long rows_per_block, int *__restrict__ err_flag, int out_vec_ok)
{
    using V = typename VecT<PW>::type;
    extern __shared__ __align__(16) float smem[];

    const int m0 = blockIdx.y * m_per_group;
    const int mg = min(m_per_group, M - m0);  // subquantizers in this group
    const int ppsq = dsub / PW;               // pieces per subquantizer
    const int ppr = mg * ppsq;                // pieces per row (this group)
    const float *qg = quantizers + (size_t)m0 * k * dsub;

    if constexpr (SMEM_CB) {
        const size_t total = (size_t)mg * k * dsub;
        if ((reinterpret_cast<uintptr_t>(qg) & 15) == 0 && (total & 3) == 0) {
            const float4 *src = reinterpret_cast<const float4 *>(qg);
            float4 *dst = reinterpret_cast<float4 *>(smem);
            for (size_t i = threadIdx.x; i < total / 4; i += kGatherThreads) dst[i] = __ldg(src + i);
        } else {
            for (size_t i = threadIdx.x; i < total; i += kGatherThreads) smem[i] = __ldg(qg + i);
        }
        __syncthreads();
    }
    const float *cbase = SMEM_CB ? smem : qg;

    const long long r0 = (long long)blockIdx.x * rows_per_block;
    const long long r1 = min(n, r0 + rows_per_block);
    bool bad = false;

    // Thread -> (piece position p, row lane rl): fixed for the whole strip, so the piece -> (m, t)
    // decomposition costs two integer divisions per thread, not per element.
    const int rows_par = ppr <= kGatherThreads ? kGatherThreads / ppr : 1;
    const int p_step = ppr <= kGatherThreads ? ppr : kGatherThreads;
    const int rl = ppr <= kGatherThreads ? (int)threadIdx.x / ppr : 0;
    if (rl >= rows_par) { if (bad) atomicExch(err_flag, 1); return; }

    for (int p = ppr <= kGatherThreads ? (int)threadIdx.x % ppr : (int)threadIdx.x; p < ppr; p += p_step) {
        const int ml = p / ppsq;
        const int t = (p - ml * ppsq) * PW;
        const long long code_col = (long long)(m0 + ml) * ccs;
        const long long out_col = (long long)(m0 + ml) * dsub + t;
        const float *csrc = cbase + (size_t)ml * k * dsub + t;

        // kUnroll independent rows in flight per thread, in three branch-free phases so that the code loads of
        // all rows are issued back to back (a data-dependent branch between them serialises the round trips).
        for (long long row = r0 + rl; row < r1; row += (long long)kUnroll * rows_par) {
            unsigned code[kUnroll];
            V v[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; u++) {
                const long long rr = min(row + (long long)u * rows_par, r1 - 1);  // clamp: re-reads the last row
                if constexpr (CW == 1) {  // u8 codes (the production shape): no width dispatch inside the loop
                    code[u] = reinterpret_cast<const uint8_t *>(codes)[rr * crs + code_col];
                } else {
                    code[u] = (unsigned)min(load_code(codes, code_width, rr * crs + code_col),
                                            (unsigned long long)0xffffffffu);
                }
            }
#pragma unroll
            for (int u = 0; u < kUnroll; u++) {
                bad |= code[u] >= (unsigned)k;
                const unsigned c = min(code[u], (unsigned)(k - 1));
                if constexpr (SMEM_CB) {
                    v[u] = *reinterpret_cast<const V *>(csrc + (size_t)c * dsub);
                } else {
                    v[u] = __ldg(reinterpret_cast<const V *>(csrc + (size_t)c * dsub));
                }
            }
#pragma unroll
            for (int u = 0; u < kUnroll; u++) {
                const long long rr = row + (long long)u * rows_par;
                if (rr < r1) {
                    float *dst = out + rr * ldo + out_col;
                    if (out_vec_ok) {
                        store_streaming<PW>(dst, v[u]);
                    } else {  // misaligned output view: scalar stores
                        const float *pv = reinterpret_cast<const float *>(&v[u]);
#pragma unroll
                        for (int e = 0; e < PW; e++) dst[e] = pv[e];
                    }
                }
            }
        }
        if (ppr <= kGatherThreads) break;
    }
    if (bad) atomicExch(err_flag, 1);
}

template <int PW>
rb_status launch_pw(const DeviceCodebook &cb, const void *codes, int code_width, size_t n, ptrdiff_t crs,
                    ptrdiff_t ccs, float *out, ptrdiff_t ldo, int *err_flag, cudaStream_t stream)
{
    const int M = (int)cb.M, k = (int)cb.k, dsub = (int)cb.dsub;
    // vector loads of the codebook need (code*dsub + t) * 4 aligned to PW*4: true because PW | dsub and
    // the codebook base comes from cudaMalloc; vector stores need the output base and ldo aligned too.
    const int out_vec_ok =
        ((reinterpret_cast<uintptr_t>(out) % (PW * sizeof(float))) == 0 && (ldo % PW) == 0) ? 1 : 0;

    // column groups: as many subquantizers as fit in ~100 KB of shared memory (2 blocks of 512 threads / SM)
    const size_t per_m = (size_t)k * dsub * sizeof(float);
    const size_t budget = 100 * 1024;
    int m_per_group = (int)(budget / per_m);
    const bool smem_cb = m_per_group >= 1;
    if (!smem_cb) m_per_group = M;
    if (m_per_group > M) m_per_group = M;
    // balance the groups
    const int n_groups = (int)ceil_div(M, m_per_group);
    m_per_group = (int)ceil_div(M, n_groups);
    const size_t smem = smem_cb ? (size_t)m_per_group * per_m : 0;

    // row strips: 2 waves of 2 blocks/SM over all groups, at least 64 rows each
    size_t strips = ceil_div((size_t)148 * 4, (size_t)n_groups);
    size_t rows_per_block = ceil_div(n, strips);
    if (rows_per_block < 64) rows_per_block = 64;
    strips = ceil_div(n, rows_per_block);
    dim3 grid((unsigned)strips, (unsigned)n_groups);

    if (smem_cb) {
        auto kern = code_width == 1 ? gather_kernel<PW, true, 1> : gather_kernel<PW, true, 0>;
        if (smem > 48 * 1024)
            RB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kGatherThreads, smem, stream>>>(cb.quantizers, k, dsub, M, codes, code_width, (long long)n,
                                                    (long long)crs, (long long)ccs, out, (long long)ldo,
                                                    m_per_group, (long long)rows_per_block, err_flag, out_vec_ok);
    } else {
        auto kern = code_width == 1 ? gather_kernel<PW, false, 1> : gather_kernel<PW, false, 0>;
        kern<<<grid, kGatherThreads, 0, stream>>>(
            cb.quantizers, k, dsub, M, codes, code_width, (long long)n, (long long)crs, (long long)ccs, out,
            (long long)ldo, m_per_group, (long long)rows_per_block, err_flag, out_vec_ok);
    }
    RB_LAUNCH_CHECK();
    return RB_OK;
}

}  // namespace

rb_status launch_gather(const DeviceCodebook &cb, const void *codes, int code_width, size_t n, ptrdiff_t crs,
                        ptrdiff_t ccs, float *out, ptrdiff_t ldo, int *err_flag, cudaStream_t stream)
{
    if (n == 0) return RB_OK;
    if (cb.dsub % 4 == 0) return launch_pw<4>(cb, codes, code_width, n, crs, ccs, out, ldo, err_flag, stream);
    if (cb.dsub % 2 == 0) return launch_pw<2>(cb, codes, code_width, n, crs, ccs, out, ldo, err_flag, stream);
    return launch_pw<1>(cb, codes, code_width, n, crs, ccs, out, ldo, err_flag, stream);
}

}  // namespace rb
