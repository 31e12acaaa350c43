// project_tc.cu — the OPQ rotation  y = x . R  on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// Replaces   x.dot(projection)                     src/pq/pq.rs:276   (before the argmin)
//            reconstructions.dot(&projection.t())  src/pq/pq.rs:324   (after the gather)
// for large batches.  The reference multiplies in FP32 (matrixmultiply: sequential FMA chain per output element);
// project.cu reproduces that order bit for bit on the FP32 pipe and is bound by it (48.8 TFLOP/s measured).  This
// kernel trades the last bits for the tensor pipe:
//   * both operands are split into two FP16 limbs of a power-of-two-scaled value (x = xh + xl, r = rh + rl: 22
//     significant bits each) and the three products xh.rh + xh.rl + xl.rh are accumulated in FP32 in tensor memory —
//     |y~ - y| <= kappa * sum_i |x_i r_ij| with kappa ~ 2^-15 worst case (dominated by the accumulation order, as for
//     any FP32 GEMM) and ~ 2^-22 typically;
//   * decode (reconstruct_batch of an Opq quantizer) uses y~ directly: north_star's tolerance for the rotated
//     reconstruction is 1e-5;
//   * encode uses y~ only to DECIDE codes that are not close: the per-row error bound is handed to the tensor encode
//     kernel, which widens its margin by it and re-decides every (row, subquantizer) pair inside the margin from an
//     exactly re-projected subvector (encode_exact.cu) — the emitted codes stay bit-exact.
//
// Work split: unit = (128-row tile, group of NT <= 256 output columns); persistent CTAs walk the units with the
// groups of one tile adjacent, so x is read from HBM once.  Per unit the K loop runs over chunks of 32 x-columns:
//   warp 13     x producer: one TMA tensor-map box (128 rows x 32 floats, 128-byte swizzle) per chunk
//   warp 14     B producer: the chunk's pre-split R limbs (core-matrix layout, one contiguous cp.async.bulk)
//   warps 4-11  converters (thread = row, two sets alternating over the chunks): FP32 -> two FP16 limbs, K-major
//               core-matrix A operand, per-chunk energies for the row's error bound
//   warp 12     one thread issues 6 tcgen05.mma (M128 x NT x K16) per chunk and the commits
//   warps 0-3   epilogue (thread = row = TMEM lane): accumulator -> registers -> rescale -> swizzled staging tile in
//               shared memory -> TMA tensor store (32 columns x 128 rows per store)
// Bounds (C4: 1M x 300): HBM 8*d B/row = 2.4 GB -> 0.37 ms; tensor 6*d*d' flop/row (d' = padded width) = 0.26 ms.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "project_tc.cuh"
#include "sm100_ptx.cuh"

namespace rb {

namespace {

using namespace ptx;

constexpr int kPT = 128;  // rows per tile (UMMA M)
constexpr int kKC = 32;   // x columns per chunk (two K = 16 slices)
constexpr int kXP = 32;   // floats per row of an x stage: 128 B, loaded with the 128-byte TMA swizzle -> the 16-byte chunk
                          // i of row r sits at chunk i ^ (r & 7), so the row-per-thread 128-bit reads are conflict free
constexpr int kMaxXS = 8; // x stages (an even number: see the converters)
constexpr int kAS = 2;    // A stages, one per converter set
constexpr int kMaxBS = 4; // B stages: the R limbs are re-streamed from L2 for every unit, and the depth of this ring
                          // (bytes in flight) is what the kernel's throughput follows (measured 3 -> 6 stages: +10 %)
constexpr int kPThreads = 32 * 15;
// warps 0-3 epilogue, 4-11 converters (two sets alternating over the K chunks), then MMA issuer and the two producers
constexpr int kWarpConv0 = 4, kWarpMma = 12, kWarpXProd = 13, kWarpBProd = 14;
constexpr int X_STAGE = kPT * kXP * 4;        // 16 384 B
constexpr int A_LIMB = (kKC / 8) * kPT * 16;  // 8 192 B: [4 core columns][128 rows][8 halves]
constexpr int A_STAGE = 2 * A_LIMB;
constexpr int Y_STAGE = kPT * 128;            // 16 384 B: 128 rows x 32 floats, 128-byte swizzled (TMA store source)
constexpr int kSmemLimit = 227 * 1024;
constexpr float kHalfLimit = 32768.f;  // |x * sx| must stay below this for the FP16 split

struct ProjParams {
    const unsigned char *bop;  // [n_groups][n_chunks][2 limbs][4 core columns][NT][8 halves]
    float *y;
    long long ldy;
    float *rowerr;        // optional: bound on |y~ - y_ref| per component of the row (NaN: the row cannot be handled)
    const float *chunk_w; // [n_chunks] error weight of a K chunk (see project_tensor_error_model)
    float limb_coef;
    const float *sx_dev;  // optional device-side scale (else sx_host)
    float sx_host, sr;
    int d, NT, n_groups, n_chunks, b_stages, x_stages;
    long long n, n_units;
};

__global__ void __launch_bounds__(kPThreads, 1) project_tc_kernel(const __grid_constant__ ProjParams p,
                                                                  const __grid_constant__ CUtensorMap tmap,
                                                                  const __grid_constant__ CUtensorMap ymap)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int B_STAGE = 128 * p.NT;
    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();  // the swizzle pattern below assumes this alignment
    unsigned char *sY = smem;  // two output staging tiles, 1 024-byte aligned for the 128-byte swizzle
    unsigned char *sX = sY + 2 * Y_STAGE;
    unsigned char *sA = sX + (size_t)p.x_stages * X_STAGE;
    unsigned char *sB = sA + kAS * A_STAGE;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sB + (size_t)p.b_stages * B_STAGE);
    uint64_t *x_full = bars, *x_empty = bars + 8, *a_full = bars + 16, *a_empty = bars + 20, *b_full = bars + 24,
             *b_empty = bars + 32, *acc_full = bars + 40, *acc_empty = bars + 42;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 44);
    float *sE = reinterpret_cast<float *>(bars + 48);  // [2][n_chunks][128] chunk energies (only with rowerr)

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.x_stages; i++) {
            mbar_init(&x_full[i], 1);
            mbar_init(&x_empty[i], 4);
        }
        for (int i = 0; i < kAS; i++) {
            mbar_init(&a_full[i], 4);
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < p.b_stages; i++) {
            mbar_init(&b_full[i], 1);
            mbar_init(&b_empty[i], 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 4);
        }
        fence_mbar_init();
    }
    if (warp == kWarpMma) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const float sx = p.sx_dev ? p.sx_dev[0] : p.sx_host;

    if (warp == kWarpXProd) {
        if (lane == 0) {
            prefetch_tensormap(&tmap);
            int s = 0;
            uint32_t ph = 1;
            for (long long u = blockIdx.x; u < p.n_units; u += gridDim.x) {
                const long long t = u / p.n_groups;
                for (int c = 0; c < p.n_chunks; c++) {
                    mbar_wait(&x_empty[s], ph);
                    mbar_arrive_expect_tx(&x_full[s], (uint32_t)X_STAGE);
                    // rows / columns outside the matrix arrive as zeros
                    tma_load_2d(sX + (size_t)s * X_STAGE, &tmap, c * kKC, (int)(t * kPT), &x_full[s]);
                    if (++s == p.x_stages) {
                        s = 0;
                        ph ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == kWarpBProd) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 1;
            for (long long u = blockIdx.x; u < p.n_units; u += gridDim.x) {
                const int g = (int)(u % p.n_groups);
                const unsigned char *src = p.bop + (size_t)g * p.n_chunks * B_STAGE;
                for (int c = 0; c < p.n_chunks; c++) {
                    mbar_wait(&b_empty[s], ph);
                    mbar_arrive_expect_tx(&b_full[s], (uint32_t)B_STAGE);
                    bulk_g2s(sB + (size_t)s * B_STAGE, src + (size_t)c * B_STAGE, (uint32_t)B_STAGE, &b_full[s]);
                    if (++s == p.b_stages) {
                        s = 0;
                        ph ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == kWarpMma) {
        if (lane == 0) {
            const uint32_t idesc = idesc_f16(kPT, (uint32_t)p.NT, 0);
            const uint32_t nt16 = (uint32_t)p.NT * 16;
            int as = 0, bs = 0;
            uint32_t aph = 0, bph = 0, un = 0;
            for (long long u = blockIdx.x; u < p.n_units; u += gridDim.x, un++) {
                const uint32_t buf = un & 1;
                mbar_wait(&acc_empty[buf], ((un >> 1) & 1) ^ 1);
                const uint32_t dst = tmem_base + buf * 256;
                for (int c = 0; c < p.n_chunks; c++) {
                    mbar_wait(&a_full[as], aph);
                    mbar_wait(&b_full[bs], bph);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(sA + (size_t)as * A_STAGE), b0 = smem_u32(sB + (size_t)bs * B_STAGE);
#pragma unroll
                    for (int ks = 0; ks < kKC / 16; ks++) {
                        const uint64_t ah = smem_desc_kmajor(a0 + ks * (2 * kPT * 16), kPT * 16, 128);
                        const uint64_t al = smem_desc_kmajor(a0 + A_LIMB + ks * (2 * kPT * 16), kPT * 16, 128);
                        const uint64_t bh = smem_desc_kmajor(b0 + ks * (2 * nt16), nt16, 128);
                        const uint64_t bl = smem_desc_kmajor(b0 + 4 * nt16 + ks * (2 * nt16), nt16, 128);
                        mma_f16_ss(dst, ah, bh, idesc, (c | ks) != 0 ? 1u : 0u);
                        mma_f16_ss(dst, ah, bl, idesc, 1u);
                        mma_f16_ss(dst, al, bh, idesc, 1u);
                    }
                    tc_commit(&a_empty[as]);
                    tc_commit(&b_empty[bs]);
                    if (++as == kAS) {
                        as = 0;
                        aph ^= 1;
                    }
                    if (++bs == p.b_stages) {
                        bs = 0;
                        bph ^= 1;
                    }
                }
                tc_commit(&acc_full[buf]);
            }
        }
        __syncwarp();
    } else if (warp >= kWarpConv0) {
        // ===================== converters (thread = row) =====================
        // Two sets of four warps take the chunks alternately (global chunk counter parity), so that one set's
        // convert -> publish latency overlaps the other's; thread = row within a set.  The numbers of x and A stages
        // are even, so a set always meets the same stages and sees EVERY phase of their barriers — a waiter that
        // skips a phase of an mbarrier cannot tell "two phases later" from "not yet".
        const int set = (warp - kWarpConv0) >> 2;
        const int row = ((warp - kWarpConv0) & 3) * 32 + lane;
        int xs = set, as = set;
        uint32_t xph = 0, aph = 1, it = 0, un = 0;
        for (long long u = blockIdx.x; u < p.n_units; u += gridDim.x, un++) {
            const long long t = u / p.n_groups;
            const int g = (int)(u % p.n_groups);
            const bool want_err = p.rowerr != nullptr && g == 0;
            float *se = sE + (size_t)(un & 1) * p.n_chunks * kPT;
            for (int c = 0; c < p.n_chunks; c++, it++) {
                if ((int)(it & 1) != set) continue;
                mbar_wait(&x_full[xs], xph);
                const unsigned char *xr = sX + (size_t)xs * X_STAGE + (size_t)row * kXP * 4;
                float4 v[kKC / 4];
#pragma unroll
                for (int i = 0; i < kKC / 4; i++) v[i] = *reinterpret_cast<const float4 *>(xr + ((i ^ (row & 7)) * 16));
                uint32_t hw[kKC / 2], lw[kKC / 2];
                float big = 0.f;
                unsigned long long ss2 = 0ull;  // the chunk's energy as two interleaved partial sums
                const unsigned long long sx2 = pack2(sx, sx), neg2 = pack2(-1.f, -1.f);
#pragma unroll
                for (int i = 0; i < kKC / 4; i++) {
                    const float e[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        // packed FP32 on the element pair (mul / fma .f32x2)
                        const unsigned long long a2 = pack2(e[2 * h], e[2 * h + 1]);
                        ss2 = fma2(a2, a2, ss2);
                        const unsigned long long s2 = mul2(a2, sx2);
                        const float s0 = lo2(s2), s1 = hi2(s2);
                        big = fmaxf(big, fmaxf(fabsf(s0), fabsf(s1)));
                        const __half2 hh = __floats2half2_rn(s0, s1);
                        const float2 hf = __half22float2(hh);
                        const unsigned long long l2 = fma2(pack2(hf.x, hf.y), neg2, s2);  // s - hf, one rounding
                        const __half2 ll = __floats2half2_rn(lo2(l2), hi2(l2));
                        hw[2 * i + h] = *reinterpret_cast<const uint32_t *>(&hh);
                        lw[2 * i + h] = *reinterpret_cast<const uint32_t *>(&ll);
                    }
                }
                const float ss = lo2(ss2) + hi2(ss2);
                // chunk energy; NaN marks a chunk the split cannot represent (NaN / Inf make ss itself non-finite)
                if (want_err) se[c * kPT + row] = (big < kHalfLimit && ss < 3.0e38f) ? ss : __int_as_float(0x7fc00000);
                mbar_wait(&a_empty[as], aph);
                unsigned char *a = sA + (size_t)as * A_STAGE;
#pragma unroll
                for (int cc = 0; cc < kKC / 8; cc++) {
                    *reinterpret_cast<uint4 *>(a + ((size_t)cc * kPT + row) * 16) =
                        make_uint4(hw[4 * cc], hw[4 * cc + 1], hw[4 * cc + 2], hw[4 * cc + 3]);
                    *reinterpret_cast<uint4 *>(a + A_LIMB + ((size_t)cc * kPT + row) * 16) =
                        make_uint4(lw[4 * cc], lw[4 * cc + 1], lw[4 * cc + 2], lw[4 * cc + 3]);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&a_full[as]);
                    mbar_arrive(&x_empty[xs]);  // every lane has consumed its x values by now
                }
                xs += 2;
                if (xs >= p.x_stages) {
                    xs -= p.x_stages;
                    xph ^= 1;
                }
                as += 2;
                if (as >= kAS) {
                    as -= kAS;
                    aph ^= 1;
                }
            }
            if (want_err) {
                // both sets have written their chunk energies of this unit (the buffer alternates per unit, and a set
                // cannot run more than one unit ahead of this barrier)
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const long long grow = t * kPT + row;
                if (set == 0 && grow < p.n) {
                    // every partial sum over i < I is at most ||x[0..I)|| * max_j ||r[0..I), j|| (Cauchy-Schwarz)
                    float ss = 0.f, eb = 0.f;
                    for (int c = 0; c < p.n_chunks; c++) {
                        ss += se[c * kPT + row];
                        eb = fmaf(sqrtf(ss), __ldg(p.chunk_w + c), eb);
                    }
                    p.rowerr[grow] = fmaf(p.limb_coef, sqrtf(ss), eb) * 1.001f;  // NaN if any chunk was NaN
                }
            }
        }
    } else {
        // ===================== epilogue (thread = row = TMEM lane) =====================
        // accumulator -> registers -> rescale -> swizzled staging tile -> TMA store.  (Storing from registers, one
        // 16-byte piece per lane and row, cost 32 store wavefronts per instruction and a third of the kernel.)
        const int row = warp * 32 + lane;
        const float inv = 1.0f / (sx * p.sr);  // a power of two
        uint32_t un = 0, blk = 0;
        for (long long u = blockIdx.x; u < p.n_units; u += gridDim.x, un++) {
            const long long t = u / p.n_groups;
            const int g = (int)(u % p.n_groups);
            const uint32_t buf = un & 1;
            mbar_wait(&acc_full[buf], (un >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + buf * 256;
            for (int c0 = 0; c0 < p.NT && g * p.NT + c0 < p.d; c0 += 32, blk++) {
                unsigned char *yb = sY + (size_t)(blk & 1) * Y_STAGE;
                uint32_t v[32];
                tmem_ld32(taddr + c0, v);
                // the store issued two blocks ago has finished reading this staging tile
                if (threadIdx.x == 0) bulk_wait_group_read<1>();
                asm volatile("bar.sync 2, 128;" ::: "memory");
                tmem_wait_ld(v);
#pragma unroll
                for (int i = 0; i < 8; i++)
                    *reinterpret_cast<float4 *>(yb + (size_t)row * 128 + (size_t)((i ^ (row & 7)) * 16)) =
                        make_float4(__uint_as_float(v[4 * i]) * inv, __uint_as_float(v[4 * i + 1]) * inv,
                                    __uint_as_float(v[4 * i + 2]) * inv, __uint_as_float(v[4 * i + 3]) * inv);
                fence_proxy_async_smem();
                asm volatile("bar.sync 2, 128;" ::: "memory");
                if (threadIdx.x == 0) {
                    tma_store_2d(&ymap, g * p.NT + c0, (int)(t * kPT), yb);  // rows >= n / columns >= d are clipped
                    bulk_commit_group();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        if (threadIdx.x == 0) bulk_wait_group_read<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kWarpMma) tmem_dealloc(tmem_base, 512);
}

// R (row-major [d][d], y = x . R) -> scaled FP16 limbs in the layout the kernel streams
__global__ void proj_prepare_kernel(const float *__restrict__ r, int d, float sr, int NT, int n_groups, int n_chunks,
                                    unsigned char *__restrict__ bop)
{
    const long long total = (long long)n_groups * n_chunks * 4 * NT;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int jl = (int)(idx % NT);
    const int cc = (int)((idx / NT) % 4);
    const int c = (int)((idx / NT / 4) % n_chunks);
    const int g = (int)(idx / NT / 4 / n_chunks);
    const int j = g * NT + jl, i0 = c * kKC + cc * 8;
    __half h[8], l[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
        const int i = i0 + e;
        const float v = (i < d && j < d) ? r[(size_t)i * d + j] * sr : 0.f;
        h[e] = __float2half_rn(v);
        l[e] = __float2half_rn(v - __half2float(h[e]));
    }
    const size_t stage = (size_t)128 * NT;
    unsigned char *base = bop + ((size_t)g * n_chunks + c) * stage;
    *reinterpret_cast<uint4 *>(base + ((size_t)cc * NT + jl) * 16) = *reinterpret_cast<const uint4 *>(h);
    *reinterpret_cast<uint4 *>(base + ((size_t)(4 + cc) * NT + jl) * 16) = *reinterpret_cast<const uint4 *>(l);
}

// |x|max over a strided sample of rows -> power-of-two scale that puts it in [2^10, 2^11)
__global__ void proj_sample_scale_kernel(const float *__restrict__ x, long long n, int d, long long ldx, long long row_step,
                                         unsigned *__restrict__ amax_bits, float *__restrict__ sx_out, unsigned *__restrict__ done,
                                         unsigned n_blocks)
{
    unsigned local = 0;
    for (long long r = (long long)blockIdx.x * row_step; r < n; r += (long long)gridDim.x * row_step)
        for (int c = threadIdx.x; c < d; c += blockDim.x) {
            const unsigned b = __float_as_uint(x[r * ldx + c]) & 0x7fffffffu;
            if (b < 0x7f800000u) local = max(local, b);  // non-finite values are dealt with per row
        }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) local = max(local, __shfl_xor_sync(0xffffffffu, local, off));
    if ((threadIdx.x & 31) == 0 && local) atomicMax(amax_bits, local);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(done, 1u) == n_blocks - 1) {
            const float amax = __uint_as_float(atomicMax(amax_bits, 0u));
            float s = 1.f;
            if (amax > 0.f) {
                int e = 10 - ilogbf(amax);  // amax * 2^e in [2^10, 2^11)
                e = max(-100, min(100, e));
                s = scalbnf(1.f, e);
            }
            sx_out[0] = s;
        }
    }
}

rb_status make_tensor_map(const float *x, size_t n, size_t d, ptrdiff_t ldx, CUtensorMap *out, unsigned box_cols = kXP,
                          bool swizzle128 = false)
{
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = []() -> EncodeFn {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeFn>(fn);
    }();
    if (!encode) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return RB_ERR_CUDA;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)n};
    const cuuint64_t strides[1] = {(cuuint64_t)ldx * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)kPT};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(x), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (n=%zu d=%zu ldx=%td)", (int)r, n, d, ldx);
        return RB_ERR_CUDA;
    }
    return RB_OK;
}

float pow2_scale_for(float amax, int target_exp)
{
    if (!(amax > 0.f) || !std::isfinite(amax)) return 1.f;
    int e = target_exp - std::ilogb(amax);
    if (e > 100) e = 100;
    if (e < -100) e = -100;
    return std::ldexp(1.f, e);
}

}  // namespace

bool project_tensor_shape_supported(size_t d) { return d >= 32 && d % 4 == 0 && d <= 4096; }

rb_status ProjTensorOperands::prepare(const float *r_dev, const float *r_host, size_t d_, cudaStream_t stream)
{
    release();
    if (!project_tensor_shape_supported(d_)) return RB_OK;
    d = (int)d_;
    // scale from |r|max (host copy); non-finite entries -> tensor path off
    float amax = 0.f;
    double colmax2 = 0.0;
    std::vector<double> col2(d_, 0.0);
    for (size_t i = 0; i < d_; i++)
        for (size_t j = 0; j < d_; j++) {
            const float v = r_host[i * d_ + j];
            if (!std::isfinite(v)) return RB_OK;
            amax = std::fmax(amax, std::fabs(v));
            col2[j] += (double)v * v;
        }
    for (size_t j = 0; j < d_; j++) colmax2 = std::fmax(colmax2, col2[j]);
    if (!(amax > 0.f)) return RB_OK;
    sr = pow2_scale_for(amax, 12);  // |r| * sr < 2^13
    rabsmax = amax;
    rcolmax = (float)(std::sqrt(colmax2) * (1.0 + 1e-6));
    const int dp16 = (d + 15) / 16 * 16;
    n_groups = (dp16 + 255) / 256;
    NT = ((dp16 + n_groups - 1) / n_groups + 31) / 32 * 32;  // whole 32-column output blocks per group
    n_chunks = (d + kKC - 1) / kKC;
    bytes = (size_t)n_groups * n_chunks * 128 * NT;
    // Error model of y~_j against the reference's FP32 y_j (encode only).  Per K chunk of 32:
    //   tensor accumulation   6 instructions, each <= 4 ulp = 8u of the largest partial sum so far (u = 2^-24;
    //                         measured <= 2.9 ulp, profiles/r1_microbench_probe.txt; 17 truncated addends with three
    //                         guard bits + the final truncation bound it by 3.2)
    //   reference FMA chain   32 roundings, each <= u of the partial sum (+ 2u for the kc = 256 block add at the end)
    // and every partial sum over i < I is at most ||x[0..I)|| * max_j ||r[0..I), j|| (Cauchy-Schwarz on the signed sum).
    // Operand split (x - xh - xl, r - rh - rl, dropped xl.rl) <= 3 * 2^-22 * ||x|| * rcolmax; limbs of r in the
    // FP16 subnormal range add sum_i |x_i| 2^-25 / sr <= ||x|| sqrt(d) 2^-25 / sr.
    {
        std::vector<double> pre(d_, 0.0);
        std::vector<float> w(n_chunks);
        const double u = std::ldexp(1.0, -24);
        for (int c = 0; c < n_chunks; c++) {
            const size_t i1 = std::min(d_, (size_t)(c + 1) * kKC);
            double mx = 0.0;
            for (size_t j = 0; j < d_; j++) {
                for (size_t i = (size_t)c * kKC; i < i1; i++) pre[j] += (double)r_host[i * d_ + j] * r_host[i * d_ + j];
                mx = std::fmax(mx, pre[j]);
            }
            w[c] = (float)((48.0 + 32.0 + (c == n_chunks - 1 ? 2.0 : 0.0)) * u * std::sqrt(mx) * 1.01);
        }
        limb_coef = (float)((3.0 * std::ldexp(1.0, -22) * rcolmax + std::sqrt((double)d_) * std::ldexp(1.0, -25) / sr) * 1.01);
        RB_CUDA_TRY(cudaMalloc(&chunk_w, n_chunks * sizeof(float)));
        RB_CUDA_TRY(cudaMemcpy(chunk_w, w.data(), n_chunks * sizeof(float), cudaMemcpyHostToDevice));
    }
    RB_CUDA_TRY(cudaMalloc(&bop, bytes));
    const long long total = (long long)n_groups * n_chunks * 4 * NT;
    proj_prepare_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(r_dev, d, sr, NT, n_groups, n_chunks,
                                                                            reinterpret_cast<unsigned char *>(bop));
    RB_LAUNCH_CHECK();
    return RB_OK;
}

void ProjTensorOperands::release()
{
    if (bop) cudaFree(bop);
    if (chunk_w) cudaFree(chunk_w);
    bop = nullptr;
    chunk_w = nullptr;
    bytes = 0;
}

bool project_tensor_call_supported(const ProjTensorOperands &ops, const float *x, size_t n, ptrdiff_t ldx, const float *y,
                                   ptrdiff_t ldy)
{
    if (!ops.ready() || n < 1024 || n >= ((size_t)1 << 31) - 256) return false;
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (ldx & 3) || ldx < (ptrdiff_t)ops.d) return false;
    if ((reinterpret_cast<uintptr_t>(y) & 15) || (ldy & 3) || ldy < (ptrdiff_t)ops.d) return false;
    return true;
}

// the per-row error bound keeps two buffers of per-chunk energies in shared memory
bool project_tensor_rowerr_supported(const ProjTensorOperands &ops) { return ops.ready() && ops.n_chunks <= 32; }

rb_status launch_project_sample_scale(const float *x, size_t n, size_t d, ptrdiff_t ldx, float *scratch4, cudaStream_t stream)
{
    // scratch4: [0] = sx (out), [1] = |x|max bits, [2] = block counter
    RB_CUDA_TRY(cudaMemsetAsync(scratch4, 0, 4 * sizeof(float), stream));
    const long long sample_rows = 2048;
    const long long row_step = n > (size_t)sample_rows ? (long long)(n / sample_rows) : 1;
    const unsigned blocks = 512;  // four sampled rows per block: the reads of a block are one round trip deep
    proj_sample_scale_kernel<<<blocks, 128, 0, stream>>>(x, (long long)n, (int)d, (long long)ldx, row_step,
                                                         reinterpret_cast<unsigned *>(scratch4 + 1), scratch4,
                                                         reinterpret_cast<unsigned *>(scratch4 + 2), blocks);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

float project_scale_for_absmax(float amax) { return pow2_scale_for(amax, 10); }

float project_tensor_error_floor(const ProjTensorOperands &ops)
{
    // R limbs in the FP16 subnormal range: sum_i |x_i sx| 2^-25 ... see prepare(); this is the part that scales with
    // 1 / sx (x limbs in the subnormal range): sum_i |r_ij| 2^-25 / sx <= rcolmax sqrt(d) 2^-25 / sx
    return (float)(ops.rcolmax * std::sqrt((double)ops.d) * std::ldexp(1.0, -25) * 1.01);
}

rb_status launch_project_tensor(const ProjTensorOperands &ops, const float *x, size_t n, ptrdiff_t ldx, const float *sx_dev,
                                float sx_host, float *y, ptrdiff_t ldy, float *rowerr, cudaStream_t stream)
{
    ProjParams p;
    p.bop = reinterpret_cast<const unsigned char *>(ops.bop);
    p.y = y;
    p.ldy = (long long)ldy;
    p.rowerr = rowerr;
    p.chunk_w = ops.chunk_w;
    p.limb_coef = ops.limb_coef;
    p.sx_dev = sx_dev;
    p.sx_host = sx_host;
    p.sr = ops.sr;
    p.d = ops.d;
    p.NT = ops.NT;
    p.n_groups = ops.n_groups;
    p.n_chunks = ops.n_chunks;
    p.n = (long long)n;
    p.n_units = (long long)ceil_div(n, (size_t)kPT) * ops.n_groups;
    const size_t b_stage = (size_t)128 * ops.NT;
    const size_t e_bytes = rowerr ? (size_t)2 * ops.n_chunks * kPT * sizeof(float) : 0;
    const size_t fixed = (size_t)2 * Y_STAGE + (size_t)kAS * A_STAGE + 48 * sizeof(uint64_t) + e_bytes;
    // four x stages (an even number, see the converters), then as many B stages as fit, then the rest to x
    int xs = 4;
    auto b_fit = [&](int xstages) {
        const size_t used = fixed + (size_t)xstages * X_STAGE;
        return used < (size_t)kSmemLimit ? (int)((kSmemLimit - used) / b_stage) : 0;
    };
    int bs = b_fit(xs);
    if (bs < 2) {
        xs = 2;
        bs = b_fit(xs);
    }
    if (bs < 2) {
        set_error("project_tc: no room for the operand rings (NT=%d, d=%d)", ops.NT, ops.d);
        return RB_ERR_UNSUPPORTED;
    }
    if (bs > 8) bs = 8;
    {
        int more = (int)((kSmemLimit - fixed - (size_t)bs * b_stage) / X_STAGE);
        more -= more % 2;
        xs = more > kMaxXS ? kMaxXS : more;
    }
    p.b_stages = bs;
    p.x_stages = xs;
    const size_t smem = fixed + (size_t)xs * X_STAGE + (size_t)bs * b_stage;
    CUtensorMap tmap, ymap;
    RB_TRY(make_tensor_map(x, n, (size_t)ops.d, ldx, &tmap, kXP, true));
    RB_TRY(make_tensor_map(y, n, (size_t)ops.d, ldy, &ymap, 32, true));
    RB_CUDA_TRY(cudaFuncSetAttribute(project_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long grid = sms;
    // keep the groups of one tile on adjacent CTAs at the same time: a grid that is a multiple of n_groups
    grid -= grid % ops.n_groups;
    if (grid < ops.n_groups) grid = ops.n_groups;
    if (grid > p.n_units) grid = p.n_units;
    project_tc_kernel<<<(unsigned)grid, kPThreads, smem, stream>>>(p, tmap, ymap);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

}  // namespace rb
