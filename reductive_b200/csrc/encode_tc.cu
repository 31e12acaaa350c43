// encode_tc.cu — nearest-centroid encode on the 5th-generation tensor cores (tcgen05, sm_100a), bit-exact.
//
// Replaces the same reference chain as encode_exact.cu
//   primitives::quantize_batch_into (src/pq/primitives.rs:64-104) -> kmeans::cluster_assignments
//   (src/kmeans.rs:133-159) -> SquaredEuclideanDistance<Ix2> (src/linalg.rs:150-180)
// for k = 256 centroids.  The reference's decision for one (row, subquantizer) is
//   argmin_j  d_j,   d_j = fl(fl(xs + cs_j) - fl(2 dp_j)),  dp_j a sequential FP32 FMA chain      (linalg.rs:167-176)
// which cannot be reproduced bit for bit by a tensor-core contraction.  What can be done exactly is to DECIDE the
// argmin from a high-precision approximation whenever the decision is not close:
//
//   1. Scores  s~_j = cs_j - 2 x.c_j  for all 256 centroids come from tcgen05.mma (M=128 rows, N=256, FP32
//      accumulators in tensor memory).  Operands are two-limb FP16 splits (x = xh + xl, c = ch + cl, 22 significant
//      bits each) of the power-of-two-scaled inputs; the three products xh.ch + xh.cl + xl.ch and the two-limb
//      ||c||^2 are laid along K (K = 3*dsub + 2, padded to a multiple of 16), so one accumulator holds the finished
//      score to ~2^-21 relative accuracy.  xs is the same for every j and is left out.
//   2. Each epilogue thread owns one row (one TMEM lane) and scans its 256 scores with 3-input minima along two
//      partitions of the index set — 16 blocks of 16 consecutive centroids and 16 interleaved chains (j mod 16).
//      The minimum appears once in each partition; every other chain / block minimum is the score of some other
//      centroid, so "exactly one block minimum and exactly one chain minimum lie below min + margin" proves that no
//      other centroid is within `margin` of the winner, and the two positions give its index (16*block + chain).
//   3. `margin` is twice a bound on |s~_j - (d_j - xs)| (operand split, tensor accumulation, and the reference's own
//      FP32 rounding), evaluated per row from ||x||^2 and max_j ||c_j||^2.  Rows that fail the test (true near-ties,
//      ~1e-4 of them on Gaussian data), rows with non-finite or out-of-range values and rows past a NaN are appended
//      to a list and re-decided by launch_encode_recheck with the reference's exact expression tree.
//   The emitted codes are therefore identical to the exact kernel's (and the oracle's) for every input.
//
// Data flow of one CTA (persistent, one per SM, 14 warps):
//   warp 13     producer: the group's B operands once (cp.async.bulk), then one TMA tensor-map tile load per row tile
//               (128 rows x the group's column slice, pitch an odd multiple of 16 bytes)
//   warps 8-11  converters: thread = row; split the FP32 subvector into FP16 limbs, write the A operand in the
//               no-swizzle K-major core-matrix layout, publish the row's margin
//   warp 12     one thread issues tcgen05.mma (K/16 instructions per unit) and tcgen05.commit
//   warps 0-7   epilogue, two sets of four warps alternating over the two 256-column accumulators
// A rotated input (x ~ x0 . R from project_tc.cu, template parameter ROT) widens the margin by the rotation's error
// bound and collects the undecided rows per subquantizer for an exact re-rotation (encode_tc.cuh RotatedInput).
// Bounds (C2: 2M x 300, M=30): HBM 4*d + M bytes per vector is the roofline (0.38 ms); the kernel is bound by the
// CUDA-core scan of the 128 x 256 accumulator (1 min3 per element and partition), see DESIGN.md.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "encode_tc.cuh"
#include "sm100_ptx.cuh"

namespace rb {

namespace {

using namespace ptx;

constexpr int kTile = 128;   // rows per tile (UMMA M)
constexpr int kCent = 256;   // centroid columns (UMMA N); codebooks with 64 < k < 256 are padded
constexpr float kPadScore = 32768.f;  // score of a padding column (scaled units, exact in FP16)
constexpr int kXStages = 2;
constexpr int kMargRing = 8;
constexpr int kThreads = 32 * 14;
// Warp roles: 8 epilogue warps (TMEM lane quarter = warp % 4), 4 converter warps, the MMA issuer and the TMA producer.
// Measured with the RB_TC_TRACE phase profile: the SM sub-partition that hosts the MMA-issuing warp scans ~35 % slower
// than the other three (560 vs 410 clocks per unit) and sets the pace; rotating the issue duty over converter warps or
// over two / four issuer warps did not pay off, so the single issuer stays.
constexpr int kWarpConv0 = 8, kWarpMma = 12, kWarpProducer = 13;
constexpr int kSmemLimit = 227 * 1024;

__host__ __device__ constexpr int kpad_of(int dsub) { return ((3 * dsub + 2 + 15) / 16) * 16; }

// error-bound constants (see header comment, DESIGN.md "tensor encode: margin"); generous by >= 4x
__device__ __forceinline__ float margin_of(float xs, float csmax, int dsub)
{
    // sqrt(xs * csmax) <= (xs + csmax) / 2
    const float e = (1.9073486e-6f + 3.8146973e-6f * (1.0f + (float)dsub * 0.03125f)) * (xs + csmax);
    return 2.0f * e;
}

// ---------------------------------------------------------------------------------------------------------
// operand preparation (runs when a codebook is created / after every k-means update)
// ---------------------------------------------------------------------------------------------------------
__global__ void tc_absmax_kernel(const float *__restrict__ q, size_t total, const float *__restrict__ cs, size_t M, size_t k,
                                 float *__restrict__ consts)
{
    // consts: [M] csmax | scale | scale2 | bad | absmax   (zeroed before this kernel)
    unsigned *amax = reinterpret_cast<unsigned *>(consts + M + 3);
    unsigned local = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned b = __float_as_uint(q[i]) & 0x7fffffffu;  // NaN / Inf compare above every finite value
        local = max(local, b);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) local = max(local, __shfl_xor_sync(0xffffffffu, local, off));
    if ((threadIdx.x & 31) == 0 && local) atomicMax(amax, local);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < M * k; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned b = __float_as_uint(cs[i]) & 0x7fffffffu;
        atomicMax(reinterpret_cast<unsigned *>(consts + i / k), b);
    }
}

__device__ __forceinline__ float scale_from_absmax(float amax, bool &bad)
{
    bad = !(amax < 3.0e38f);  // NaN or Inf somewhere in the codebook
    if (bad || amax == 0.f) return 1.f;
    int p = ilogbf(amax) + 1;       // amax < 2^p
    int e = 3 - p;                  // amax * 2^e in [4, 8)
    e = max(-60, min(60, e));
    return scalbnf(1.f, e);
}

template <int DSUB>
__global__ void tc_prepare_kernel(const float *__restrict__ q, const float *__restrict__ cs, int M, int k,
                                  __half *__restrict__ bop, float *__restrict__ consts)
{
    constexpr int KPAD = kpad_of(DSUB), NCH = KPAD / 8;
    bool bad;
    const float scale = scale_from_absmax(consts[M + 3], bad);
    const float scale2 = scale * scale;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        consts[M + 0] = scale;
        consts[M + 1] = scale2;
        consts[M + 2] = bad ? 1.f : 0.f;
    }
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (m, j)
    if (idx >= M * kCent) return;
    const int m = idx / kCent, j = idx % kCent;
    __half kv[KPAD];
#pragma unroll
    for (int t = 0; t < KPAD; t++) kv[t] = __float2half_rn(0.f);
    if (j >= k) {
        // k < 256: the unused columns get a zero centroid with the constant score kPadScore, far above any real
        // score of a row the tensor pass is allowed to decide (rows with large norms are decided exactly)
        kv[3 * DSUB] = __float2half_rn(kPadScore);
    } else {
    const float *c = q + ((size_t)m * k + j) * DSUB;
#pragma unroll
    for (int t = 0; t < DSUB; t++) {
        const float cv = c[t] * scale;
        const __half h = __float2half_rn(cv);
        const __half l = __float2half_rn(cv - __half2float(h));
        const __half h2 = __float2half_rn(-2.f * __half2float(h));  // exact (|cv| < 8)
        const __half l2 = __float2half_rn(-2.f * __half2float(l));
        kv[t] = h2;
        kv[DSUB + t] = l2;
        kv[2 * DSUB + t] = h2;
    }
    const float csv = cs[(size_t)m * k + j] * scale2;
    const __half ch = __float2half_rn(csv);
    kv[3 * DSUB] = ch;
    kv[3 * DSUB + 1] = __float2half_rn(csv - __half2float(ch));
    }
    // [m][chunk][j][8 halves]
    __half *dst = bop + (size_t)m * NCH * kCent * 8;
#pragma unroll
    for (int ch8 = 0; ch8 < NCH; ch8++) {
        uint4 w;
        __half2 *hw = reinterpret_cast<__half2 *>(&w);
#pragma unroll
        for (int e = 0; e < 4; e++) hw[e] = __halves2half2(kv[ch8 * 8 + 2 * e], kv[ch8 * 8 + 2 * e + 1]);
        *reinterpret_cast<uint4 *>(dst + ((size_t)ch8 * kCent + j) * 8) = w;
    }
}

// ---------------------------------------------------------------------------------------------------------
// the encode kernel
// ---------------------------------------------------------------------------------------------------------
constexpr int kMaxGroups = 148;

struct EncParams {
    const float *x;
    long long n;
    const __half *bop;
    const float *consts;  // [M] csmax | scale | scale2 | bad | absmax
    void *codes;
    int code_width;
    long long crs, ccs;
    uint32_t *pairs;
    uint32_t *n_pairs;
    uint32_t max_pairs;
    int M, gm, n_groups, a_stages, pitch_f;
    float xs_limit;  // rows with ||x * scale||^2 at or above this are decided exactly
    // x is an APPROXIMATE rotation of other rows (project_tc.cu): every component of a row may be off by
    // rowerr[row] + perr_floor / *perr_sx, which widens the margin; rowerr == nullptr: x is exact
    uint32_t *bucket_counts, *bucket_rows;  // rotated input: flagged rows per subquantizer, [M] and [M][n]
    const float *rowerr;
    const float *perr_sx;  // device scalar: the rotation's operand scale
    float perr_floor;
    long long n_tiles;
    long long *trace;  // debugging aid (RB_TC_TRACE): per-role clock64 stamps of CTA 0, else nullptr
    unsigned short cta_start[kMaxGroups + 1];  // CTAs [cta_start[g], cta_start[g+1]) own column group g
};

__device__ __forceinline__ float fmin3(float a, float b, float c)
{
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float fma_sat(float a, float b, float c)
{
    float r;
    asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float min16(const uint32_t *v)
{
    const float t0 = fmin3(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]));
    const float t1 = fmin3(__uint_as_float(v[3]), __uint_as_float(v[4]), __uint_as_float(v[5]));
    const float t2 = fmin3(__uint_as_float(v[6]), __uint_as_float(v[7]), __uint_as_float(v[8]));
    const float t3 = fmin3(__uint_as_float(v[9]), __uint_as_float(v[10]), __uint_as_float(v[11]));
    const float t4 = fmin3(__uint_as_float(v[12]), __uint_as_float(v[13]), __uint_as_float(v[14]));
    const float r0 = fmin3(t0, t1, t2);
    const float r1 = fmin3(t3, t4, __uint_as_float(v[15]));
    return fminf(r0, r1);
}

// Phase profile (build with -DRB_TC_PHASES, run with RB_TC_TRACE=1): lane 0 of every warp of CTA 0 accumulates clock()
// deltas per phase in registers and writes them once at the end: trace[warp * 8 + phase].  Phase meanings are per
// role (see the host-side dump).  Compiled out by default.
#ifdef RB_TC_PHASES
#define RB_PH_BEGIN() unsigned ph_t = clock(); unsigned long long ph_acc[6] = {0, 0, 0, 0, 0, 0}; (void)ph_t; (void)ph_acc
#define RB_PH(i)                                   \
    do {                                           \
        if (p.trace != nullptr) {                  \
            const unsigned now_ = clock();         \
            ph_acc[i] += (unsigned)(now_ - ph_t);  \
            ph_t = now_;                           \
        }                                          \
    } while (0)
#define RB_PH_END()                                                                                       \
    do {                                                                                                  \
        if (p.trace != nullptr && blockIdx.x == 0 && lane == 0)                                           \
            for (int i_ = 0; i_ < 6; i_++) p.trace[warp * 8 + i_] = (long long)ph_acc[i_];                \
    } while (0)
#else
#define RB_PH_BEGIN() do { } while (0)
#define RB_PH(i) do { } while (0)
#define RB_PH_END() do { } while (0)
#endif

// ROT: x is an approximate rotation (EncParams::rowerr); a separate instantiation so that the plain kernel carries
// none of its code (the run-time branch cost 3 % on C2)
template <int DSUB, bool ROT>
__global__ void __launch_bounds__(kThreads, 1) encode_tc_kernel(const __grid_constant__ EncParams p,
                                                                const __grid_constant__ CUtensorMap tmap)
{
    static_assert(DSUB % 2 == 0, "the converter packs pairs of elements");
    constexpr int KPAD = kpad_of(DSUB), NCH = KPAD / 8;
    constexpr int A_BYTES = NCH * kTile * 16, B_BYTES = NCH * kCent * 16;
    extern __shared__ __align__(128) unsigned char smem[];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int xs_bytes = kTile * p.pitch_f * 4;  // one X stage: 128 rows x pitch
    unsigned char *sB = smem;
    unsigned char *sX = sB + (size_t)p.gm * B_BYTES;
    unsigned char *sA = sX + (size_t)kXStages * xs_bytes;
    float *sMarg = reinterpret_cast<float *>(sA + (size_t)p.a_stages * A_BYTES);  // [kMargRing][kTile]
    // rotated input: two more per-row values (margin = sMarg + sMargB * sqrt(max(sMargC + min score, 0)), see below)
    float *sMargB = sMarg + kMargRing * kTile, *sMargC = sMargB + kMargRing * kTile;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sMarg + (ROT ? 3 : 1) * kMargRing * kTile);
    uint64_t *x_full = bars, *x_empty = bars + 2, *a_full = bars + 4, *a_empty = bars + 8, *acc_full = bars + 12,
             *acc_empty = bars + 14, *b_full = bars + 16;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 20);

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(&x_full[i], 1);
            mbar_init(&x_empty[i], 4);
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 4);
        }
        for (int i = 0; i < 4; i++) {
            mbar_init(&a_full[i], 4);
            mbar_init(&a_empty[i], 1);
        }
        mbar_init(b_full, 1);
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

    // this CTA's column group g, its rank among the group's CTAs and their number: it encodes the 128-row tiles
    // rank, rank + stride, ... for the subquantizers [g*gm, g*gm + gm_cur).  The groups of one tile are handled by
    // different CTAs at about the same time, so partially used 128-byte lines of x are fetched from HBM once.
    int g = 0;
    while (g + 1 < p.n_groups && (int)blockIdx.x >= (int)p.cta_start[g + 1]) g++;
    const long long t_first = (long long)blockIdx.x - p.cta_start[g];
    const long long t_stride = (long long)p.cta_start[g + 1] - p.cta_start[g];
    const int gm_cur = min(p.gm, p.M - g * p.gm);
    const int S = p.a_stages;

    if (warp == kWarpProducer) {
        // ===================== producer =====================
        if (lane == 0) {
            prefetch_tensormap(&tmap);
            mbar_arrive_expect_tx(b_full, (uint32_t)gm_cur * B_BYTES);
            for (int ml = 0; ml < gm_cur; ml++)
                bulk_g2s(sB + (size_t)ml * B_BYTES, p.bop + (size_t)(g * p.gm + ml) * (B_BYTES / 2), B_BYTES, b_full);
            uint32_t li = 0;
            for (long long t = t_first; t < p.n_tiles; t += t_stride, li++) {
                const int stage = (int)(li & 1);
                mbar_wait(&x_empty[stage], ((li >> 1) & 1) ^ 1);
                // one TMA tile load: box = 128 rows x pitch floats; rows / columns outside the matrix arrive as zeros
                mbar_arrive_expect_tx(&x_full[stage], (uint32_t)xs_bytes);
                tma_load_2d(sX + (size_t)stage * xs_bytes, &tmap, g * p.gm * DSUB, (int)(t * kTile), &x_full[stage]);
            }
        }
        __syncwarp();
    } else if (warp == kWarpMma) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            // This single thread sits on the critical path between an accumulator being released and the next
            // scan starting, so its per-unit instruction stream is kept minimal: the shared-memory descriptors
            // differ only in the 14-bit start-address field of their low word, which is advanced by adds.
            const uint32_t idesc = idesc_f16(kTile, kCent, 0);
            const uint64_t a_desc0 = smem_desc_kmajor(smem_u32(sA), kTile * 16, 128);
            const uint64_t b_desc0 = smem_desc_kmajor(smem_u32(sB), kCent * 16, 128);
            const uint32_t a_hi = (uint32_t)(a_desc0 >> 32), b_hi = (uint32_t)(b_desc0 >> 32);
            const uint32_t a_lo0 = (uint32_t)a_desc0, b_lo0 = (uint32_t)b_desc0;
            constexpr uint32_t A_STEP = A_BYTES >> 4, B_STEP = B_BYTES >> 4;            // per stage / per subquantizer
            constexpr uint32_t A_KS = (2 * kTile * 16) >> 4, B_KS = (2 * kCent * 16) >> 4;  // per K = 16 slice
            uint32_t u = 0, as = 0, aph = 0, a_lo = a_lo0;
            RB_PH_BEGIN();
            mbar_wait(b_full, 0);
            for (long long t = t_first; t < p.n_tiles; t += t_stride) {
                uint32_t b_lo = b_lo0;
                for (int ml = 0; ml < gm_cur; ml++, u++, b_lo += B_STEP) {
                    const uint32_t buf = u & 1;
                    RB_PH(0);
                    mbar_wait(&acc_empty[buf], ((u >> 1) & 1) ^ 1);
                    RB_PH(1);
                    mbar_wait(&a_full[as], aph);
                    RB_PH(2);
                    tc_fence_after();
#pragma unroll
                    for (int ks = 0; ks < KPAD / 16; ks++)
                        mma_f16_ss_lohi(tmem_base + buf * kCent, a_lo + ks * A_KS, a_hi, b_lo + ks * B_KS, b_hi, idesc,
                                        ks > 0 ? 1u : 0u);
                    tc_commit(&acc_full[buf]);
                    tc_commit(&a_empty[as]);
                    RB_PH(3);
                    a_lo += A_STEP;
                    if (++as == (uint32_t)S) {
                        as = 0;
                        aph ^= 1;
                        a_lo = a_lo0;
                    }
                }
            }
            RB_PH_END();
        }
        __syncwarp();
    } else if (warp >= kWarpConv0) {
        // ===================== converters (thread = row) =====================
        const int row = (warp - kWarpConv0) * 32 + lane;
        const float scale = p.consts[p.M + 0];
        const float scale2 = p.consts[p.M + 1];
        const bool cb_bad = p.consts[p.M + 2] != 0.f;
        uint32_t u = 0, as = 0, aph = 0, li = 0;
        RB_PH_BEGIN();
        for (long long t = t_first; t < p.n_tiles; t += t_stride, li++) {
            const int stage = (int)(li & 1);
            RB_PH(0);
            mbar_wait(&x_full[stage], (li >> 1) & 1);
            RB_PH(1);
            const float4 *xr = reinterpret_cast<const float4 *>(sX + (size_t)stage * xs_bytes + (size_t)row * p.pitch_f * 4);
            float perr = 0.f;  // bound on the error of each component of this row (0: the row is exact)
            if constexpr (ROT) {
                const long long grow = t * kTile + row;
                perr = (grow < p.n ? p.rowerr[grow] : 0.f) + p.perr_floor / p.perr_sx[0];
            }
            // two subquantizers at a time: 2*DSUB floats are a whole number of 16-byte vectors, and with a row pitch
            // that is an odd multiple of 16 bytes the 128-bit loads of a warp are bank-conflict free
            for (int ml0 = 0; ml0 < gm_cur; ml0 += 2) {
                float xv[2 * DSUB];
#pragma unroll
                for (int q4 = 0; q4 < DSUB / 2; q4++) {
                    const float4 v = xr[(ml0 * DSUB) / 4 + q4];
                    xv[4 * q4] = v.x;
                    xv[4 * q4 + 1] = v.y;
                    xv[4 * q4 + 2] = v.z;
                    xv[4 * q4 + 3] = v.w;
                }
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    if (ml0 + h >= gm_cur) break;
                    const float *sub = xv + h * DSUB;
                    unsigned long long xs2 = 0ull;  // ||x||^2 as two interleaved partial sums (only the margin uses it)
                    uint32_t hw[DSUB / 2], lw[DSUB / 2];
#pragma unroll
                    for (int t2 = 0; t2 < DSUB / 2; t2++) {
                        const float a0 = sub[2 * t2], a1 = sub[2 * t2 + 1];
                        xs2 = fma2(pack2(a0, a1), pack2(a0, a1), xs2);
                        // packed FP32 (mul / fma .f32x2) on the element pair; two-limb FP16 split, two elements per
                        // conversion instruction
                        const unsigned long long s2 = mul2(pack2(a0, a1), pack2(scale, scale));
                        const float s0 = lo2(s2), s1 = hi2(s2);
                        const __half2 hh = __floats2half2_rn(s0, s1);
                        const float2 hf = __half22float2(hh);
                        const unsigned long long l2 = fma2(pack2(hf.x, hf.y), pack2(-1.f, -1.f), s2);  // s - hf, one rounding
                        const __half2 ll = __floats2half2_rn(lo2(l2), hi2(l2));
                        hw[t2] = *reinterpret_cast<const uint32_t *>(&hh);
                        lw[t2] = *reinterpret_cast<const uint32_t *>(&ll);
                    }
                    // K layout [xh | xh | xl | 1 1 | 0 ...] as 32-bit words (two halves each)
                    uint32_t w[KPAD / 2];
#pragma unroll
                    for (int i = 0; i < KPAD / 2; i++) w[i] = 0u;
#pragma unroll
                    for (int i = 0; i < DSUB / 2; i++) {
                        w[i] = hw[i];
                        w[DSUB / 2 + i] = hw[i];
                        w[DSUB + i] = lw[i];
                    }
                    w[3 * DSUB / 2] = 0x3c003c00u;  // (1.0, 1.0)
                    // NaN in x makes xs NaN, Inf makes it Inf; |x * scale| <= sqrt(xs) * scale must stay inside the
                    // FP16 range (2^15) and, for padded codebooks, real scores must stay far below kPadScore
                    // (cs + 2 |x||c| <= 64 + 2 * 8 * sqrt(xs_sc) < 16 384 for xs_sc < 10^6): decide other rows exactly
                    const float xs = lo2(xs2) + hi2(xs2);
                    const float xs_sc = xs * scale2;
                    const bool bad = cb_bad || !(xs_sc < p.xs_limit);
                    const float csmax = p.consts[g * p.gm + ml0 + h];
                    float marg = margin_of(xs, csmax, DSUB) * scale2;
                    float margb = 0.f, margc = 0.f;
                    if constexpr (ROT) {
                        // Rotated input: the subvector is off by delta, ||delta|| <= sqrt(dsub) * perr, which moves the
                        // DIFFERENCE of two scores by 2 delta.(c_j - c_i) <= 2 ||delta|| (a + b), a and b the true
                        // distances to the two centroids.  A centroid with b - a > 4 ||delta|| + 2 sqrt(e) (e = half the
                        // base margin) is ranked correctly by the reference whatever we measure, since then
                        // b^2 - a^2 >= (b - a)^2 > 4 e; for the others a + b <= 2 a + 4 ||delta|| + 2 sqrt(e), so
                        //   margin = base + 4 ||delta|| (a + 3 ||delta|| + sqrt(e)),  a <= sqrt(d~_min + e) + ||delta||
                        // with d~_min = xs + (minimum score), known only in the epilogue.  All in scaled units.
                        const float ds = sqrtf((float)DSUB) * perr * scale * 1.02f;
                        const float eb = 0.5f * marg;
                        marg += 4.0f * ds * (4.0f * ds + sqrtf(eb));
                        margb = 4.0f * ds;
                        margc = xs_sc + eb;
                    }
                    if (bad || !(marg < 3.0e38f) || !(margb < 3.0e38f))
                        marg = __int_as_float(0x7fc00000);  // NaN: always re-decide exactly
                    RB_PH(2);
                    mbar_wait(&a_empty[as], aph ^ 1);
                    RB_PH(3);
                    unsigned char *a = sA + (size_t)as * A_BYTES;
#pragma unroll
                    for (int c8 = 0; c8 < NCH; c8++)
                        *reinterpret_cast<uint4 *>(a + ((size_t)c8 * kTile + row) * 16) =
                            make_uint4(w[4 * c8], w[4 * c8 + 1], w[4 * c8 + 2], w[4 * c8 + 3]);
                    sMarg[(u % kMargRing) * kTile + row] = marg;
                    if constexpr (ROT) {
                        sMargB[(u % kMargRing) * kTile + row] = margb;
                        sMargC[(u % kMargRing) * kTile + row] = margc;
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&a_full[as]);
                    RB_PH(4);
                    if (++as == (uint32_t)S) {
                        as = 0;
                        aph ^= 1;
                    }
                    u++;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&x_empty[stage]);
        }
        RB_PH_END();
    } else {
        // ===================== epilogue (thread = row = TMEM lane) =====================
        const int set = warp >> 2;
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)set * kCent;
        const float INF = __int_as_float(0x7f800000);
        // this set's units are u = set, set + 2, ...: (tile, ml) advance by two subquantizers with wrap-around
        long long t = t_first;
        int ml = set;
        while (ml >= gm_cur && t < p.n_tiles) {
            ml -= gm_cur;
            t += t_stride;
        }
        RB_PH_BEGIN();
        for (uint32_t u = (uint32_t)set; t < p.n_tiles; u += 2) {
            const long long grow = t * kTile + row;
            const int m = g * p.gm + ml;
            {
                RB_PH(0);
                mbar_wait(&acc_full[set], (u >> 1) & 1);
                RB_PH(1);
                tc_fence_after();
                float A[16], B[16];
#pragma unroll
                for (int i = 0; i < 16; i++) A[i] = INF;
                uint32_t v0[32], v1[32];
                tmem_ld32(taddr, v0);
#pragma unroll
                for (int c0 = 0; c0 < kCent; c0 += 64) {
                    tmem_wait_ld(v0);
                    tmem_ld32(taddr + c0 + 32, v1);
                    B[c0 / 16] = min16(v0);
                    B[c0 / 16 + 1] = min16(v0 + 16);
#pragma unroll
                    for (int a = 0; a < 16; a++) A[a] = fmin3(A[a], __uint_as_float(v0[a]), __uint_as_float(v0[16 + a]));
                    tmem_wait_ld(v1);
                    if (c0 + 64 < kCent) tmem_ld32(taddr + c0 + 64, v0);
                    B[c0 / 16 + 2] = min16(v1);
                    B[c0 / 16 + 3] = min16(v1 + 16);
#pragma unroll
                    for (int a = 0; a < 16; a++) A[a] = fmin3(A[a], __uint_as_float(v1[a]), __uint_as_float(v1[16 + a]));
                }
                // the accumulator is free again
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[set]);
                RB_PH(2);
                float marg = sMarg[(u % kMargRing) * kTile + row];
                float margb = 0.f, margc = 0.f;
                if constexpr (ROT) {
                    margb = sMargB[(u % kMargRing) * kTile + row];
                    margc = sMargC[(u % kMargRing) * kTile + row];
                }

                // minimum of the 16 block minima as a tree (short dependency chain)
                const float ma = fmin3(B[0], B[1], B[2]), mb = fmin3(B[3], B[4], B[5]), mc = fmin3(B[6], B[7], B[8]);
                const float md = fmin3(B[9], B[10], B[11]), me = fmin3(B[12], B[13], B[14]);
                const float m1 = fminf(fmin3(ma, mb, mc), fmin3(md, me, B[15]));
                // Which block / chain minima lie below thr = m1 + margin (FMA pipe): t = sat((thr - v) * 2^40) is
                // exactly 1 for v < thr and exactly 0 for v >= thr or NaN as long as |thr| >= 2^-14 (then thr - v
                // is zero or at least ulp(thr) >= 2^-37).  acc = sum t_i * (64 + i) lies in [64, 80) iff exactly
                // one t_i is set, and then names it.  Four partial sums each keep the dependency chains short.
                if constexpr (ROT) marg = fmaf(margb, sqrtf(fmaxf(margc + m1, 0.f)), marg);
                const float thr = m1 + marg;
                const float SC = 1.099511627776e12f;  // 2^40
                const float thr_sc = thr * SC;
                // packed FP32 FMA (fma.rn.f32x2): lanes = two consecutive positions, the same weight pairs serve both
                // partitions; four partial sums keep the dependency chains short
                unsigned long long bs2[2] = {0ull, 0ull}, as2[2] = {0ull, 0ull};
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const unsigned long long w2 = pack2((float)(64 + i), (float)(65 + i));
                    bs2[(i >> 1) & 1] = fma2(pack2(fma_sat(B[i], -SC, thr_sc), fma_sat(B[i + 1], -SC, thr_sc)), w2, bs2[(i >> 1) & 1]);
                    as2[(i >> 1) & 1] = fma2(pack2(fma_sat(A[i], -SC, thr_sc), fma_sat(A[i + 1], -SC, thr_sc)), w2, as2[(i >> 1) & 1]);
                }
                const float accb = (lo2(bs2[0]) + hi2(bs2[0])) + (lo2(bs2[1]) + hi2(bs2[1]));
                const float acca = (lo2(as2[0]) + hi2(as2[0])) + (lo2(as2[1]) + hi2(as2[1]));

                // one block and one chain below thr; |thr| large enough for exact t; NaN margins fail the comparisons
                const bool certain = (fminf(accb, acca) >= 64.f) && (fmaxf(accb, acca) < 80.f) &&
                                     (fabsf(thr) >= 6.103515625e-5f) && (fabsf(m1) < 3.0e38f);
                if (grow < p.n) {
                    const unsigned code = certain ? (unsigned)(int)fmaf(accb - 64.f, 16.f, acca - 64.f) : 0u;
                    store_code(p.codes, p.code_width, grow * p.crs + (long long)m * p.ccs, code);
                    if (!certain) {
                        if (p.bucket_rows != nullptr) {  // a row is flagged at most once per subquantizer: slot < n
                            const uint32_t slot = atomicAdd(&p.bucket_counts[m], 1u);
                            p.bucket_rows[(size_t)m * (size_t)p.n + slot] = (uint32_t)grow;
                        } else {
                            const uint32_t slot = atomicAdd(p.n_pairs, 1u);
                            if (slot < p.max_pairs) {
                                p.pairs[2 * (size_t)slot] = (uint32_t)grow;
                                p.pairs[2 * (size_t)slot + 1] = (uint32_t)m;
                            }
                        }
                    }
                }
                RB_PH(3);
            }
            ml += 2;
            while (ml >= gm_cur && t < p.n_tiles) {
                ml -= gm_cur;
                t += t_stride;
            }
        }
        RB_PH_END();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kWarpMma) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
struct Plan {
    int gm = 0, n_groups = 0, a_stages = 0, pitch_f = 0, ctas = 0;
    size_t smem = 0;
    unsigned short cta_start[kMaxGroups + 1] = {0};
};

// Column grouping and CTA allocation.  A CTA keeps the B operands of one group of gm subquantizers in shared
// memory for its whole life and walks 128-row tiles; groups get CTAs in proportion to their width.  Chosen to
// minimise the busiest CTA's number of (tile, subquantizer) units.
Plan make_plan_uncached(size_t M, size_t dsub, size_t n_tiles, int sms, int marg_rings);

// The search below costs 40-250 us of host time (it is quadratic in the number of column groups), which shows in
// loops of short calls (k-means on small row counts): the last few shapes are remembered.
Plan make_plan(size_t M, size_t dsub, size_t n_tiles, int sms, int marg_rings = 1)
{
    struct Entry {
        size_t M, dsub, n_tiles;
        int sms, marg_rings;
        Plan plan;
    };
    static std::mutex mu;
    static std::vector<Entry> cache;
    {
        std::lock_guard<std::mutex> lock(mu);
        for (const Entry &e : cache)
            if (e.M == M && e.dsub == dsub && e.n_tiles == n_tiles && e.sms == sms && e.marg_rings == marg_rings) return e.plan;
    }
    const Plan plan = make_plan_uncached(M, dsub, n_tiles, sms, marg_rings);
    std::lock_guard<std::mutex> lock(mu);
    if (cache.size() >= 32) cache.erase(cache.begin());
    cache.push_back(Entry{M, dsub, n_tiles, sms, marg_rings, plan});
    return plan;
}

Plan make_plan_uncached(size_t M, size_t dsub, size_t n_tiles, int sms, int marg_rings)
{
    Plan best;
    unsigned long long best_span = ~0ull;
    const size_t kpad = (size_t)kpad_of((int)dsub), nch = kpad / 8;
    const size_t a_bytes = nch * kTile * 16, b_bytes = nch * kCent * 16;
    if ((M * dsub) % 4 != 0 || (dsub & 1) || n_tiles == 0) return best;
    for (size_t gm = M < 16 ? M : 16; gm >= 1; gm--) {
        const size_t n_groups = ceil_div(M, gm);
        if (n_groups > (size_t)kMaxGroups || n_groups > (size_t)sms) continue;
        if ((gm * dsub) % 4 != 0 && n_groups > 1) continue;  // TMA box start: 16-byte aligned column offset
        // smem row pitch: room for ceil(gm/2) pairs of subvectors, an odd number of 16-byte units
        size_t pitch_b = ((2 * dsub * ceil_div(gm, 2) * 4 + 15) / 16) * 16;
        if ((pitch_b / 16) % 2 == 0) pitch_b += 16;
        if (pitch_b / 4 > 256) continue;  // TMA box extent
        int stages = 0;
        size_t smem = 0;
        for (int s = 4; s >= 2 && !stages; s--) {
            smem = gm * b_bytes + (size_t)kXStages * kTile * pitch_b + (size_t)s * a_bytes + (size_t)marg_rings * kMargRing * kTile * 4 + 24 * 8;
            if (smem <= (size_t)kSmemLimit - 1024) stages = s;
        }
        if (!stages) continue;
        // greedy CTA allocation: every group gets one, the rest go to whichever group is busiest
        size_t ctas = (size_t)sms < n_tiles * n_groups ? (size_t)sms : n_tiles * n_groups;
        unsigned cnt[kMaxGroups];
        for (size_t g = 0; g < n_groups; g++) cnt[g] = 1;
        auto width = [&](size_t g) { return g + 1 < n_groups ? gm : M - (n_groups - 1) * gm; };
        auto load = [&](size_t g) { return (unsigned long long)ceil_div(n_tiles, (size_t)cnt[g]) * width(g); };
        for (size_t used = n_groups; used < ctas; used++) {
            size_t arg = 0;
            for (size_t g = 1; g < n_groups; g++)
                if (load(g) > load(arg)) arg = g;
            if (cnt[arg] >= n_tiles) break;
            cnt[arg]++;
        }
        unsigned long long span = 0;
        for (size_t g = 0; g < n_groups; g++) span = load(g) > span ? load(g) : span;
        if (span < best_span) {
            best_span = span;
            best.gm = (int)gm;
            best.n_groups = (int)n_groups;
            best.a_stages = stages;
            best.pitch_f = (int)(pitch_b / 4);
            best.smem = smem;
            unsigned acc = 0;
            for (size_t g = 0; g < n_groups; g++) {
                best.cta_start[g] = (unsigned short)acc;
                acc += cnt[g];
            }
            best.cta_start[n_groups] = (unsigned short)acc;
            best.ctas = (int)acc;
        }
    }
    return best;
}

// 2-D tensor map over the row-major batch: dimension 0 = the d columns, dimension 1 = the n rows (pitch ldx floats);
// box = one tile's column slice.  cuTensorMapEncodeTiled is resolved through the runtime (no libcuda link).
rb_status make_x_tensor_map(const float *x, size_t n, size_t d, ptrdiff_t ldx, size_t box_cols, CUtensorMap *out)
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
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)kTile};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(x), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (n=%zu d=%zu ldx=%td box=%zu)", (int)r, n, d, ldx, box_cols);
        return RB_ERR_CUDA;
    }
    return RB_OK;
}

int device_sm_count()
{
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms > kMaxGroups ? kMaxGroups : sms;
}

template <int DSUB>
rb_status launch_t(const DeviceCodebook &cb, const TensorOperands &tc, const float *x, size_t n, ptrdiff_t ldx, void *codes,
                   int code_width, ptrdiff_t crs, ptrdiff_t ccs, uint32_t *pairs, uint32_t *n_pairs, uint32_t max_pairs,
                   const RotatedInput *rot, cudaStream_t stream)
{
    const size_t n_tiles = ceil_div(n, (size_t)kTile);
    const Plan plan = make_plan(cb.M, cb.dsub, n_tiles, device_sm_count(), rot ? 3 : 1);
    EncParams p;
    p.x = x;
    p.n = (long long)n;
    p.bop = reinterpret_cast<const __half *>(tc.b_tiles);
    p.consts = tc.consts;
    p.codes = codes;
    p.code_width = code_width;
    p.crs = (long long)crs;
    p.ccs = (long long)ccs;
    p.pairs = pairs;
    p.n_pairs = n_pairs;
    p.max_pairs = max_pairs;
    p.M = (int)cb.M;
    p.gm = plan.gm;
    p.n_groups = plan.n_groups;
    p.a_stages = plan.a_stages;
    p.pitch_f = plan.pitch_f;
    p.n_tiles = (long long)n_tiles;
    p.xs_limit = cb.k < (size_t)kCent ? 1.0e6f : 1.0e9f;
    p.trace = nullptr;
    p.bucket_counts = rot ? n_pairs : nullptr;
    p.bucket_rows = rot ? pairs : nullptr;
    p.rowerr = rot ? rot->rowerr : nullptr;
    p.perr_sx = rot ? rot->sx_dev : nullptr;
    p.perr_floor = rot ? rot->err_floor : 0.f;
    for (int g = 0; g <= kMaxGroups; g++) p.cta_start[g] = plan.cta_start[g < plan.n_groups ? g : plan.n_groups];
    CUtensorMap tmap;
    RB_TRY(make_x_tensor_map(x, n, cb.M * cb.dsub, ldx, (size_t)plan.pitch_f, &tmap));
    auto kern = rot ? encode_tc_kernel<DSUB, true> : encode_tc_kernel<DSUB, false>;
    RB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
#ifdef RB_TC_PHASES
    const bool trace = getenv("RB_TC_TRACE") != nullptr;
#else
    const bool trace = false;
#endif
    const size_t trace_len = (size_t)(kThreads / 32) * 8;
    if (trace) {
        RB_CUDA_TRY(cudaMalloc(&p.trace, trace_len * sizeof(long long)));
        RB_CUDA_TRY(cudaMemset(p.trace, 0, trace_len * sizeof(long long)));
    }
    kern<<<(unsigned)plan.ctas, kThreads, plan.smem, stream>>>(p, tmap);
    RB_LAUNCH_CHECK();
    if (trace) {  // debugging aid: where CTA 0's warps spent their clocks, per (tile, subquantizer) unit
        std::vector<long long> h(trace_len);
        RB_CUDA_TRY(cudaStreamSynchronize(stream));
        RB_CUDA_TRY(cudaMemcpy(h.data(), p.trace, trace_len * sizeof(long long), cudaMemcpyDeviceToHost));
        cudaFree(p.trace);
        const double units = (double)ceil_div(n_tiles, (size_t)(plan.cta_start[1] - plan.cta_start[0])) *
                             (double)(plan.n_groups > 1 ? plan.gm : (int)cb.M);
        fprintf(stderr, "[rb tc phases] clocks per unit of CTA 0 (%.0f units)\n", units);
        fprintf(stderr, "  epilogue: loop | wait acc_full | scan | certificate+store\n");
        fprintf(stderr, "  converter: loop | wait x_full | load+convert | wait a_empty | write+publish\n");
        fprintf(stderr, "  mma: loop | wait acc_empty | wait a_full | issue+commit\n");
        for (int w = 0; w < kThreads / 32; w++) {
            fprintf(stderr, "  warp %2d:", w);
            for (int i = 0; i < 6; i++) fprintf(stderr, " %8.1f", (double)h[(size_t)w * 8 + i] / units);
            fprintf(stderr, "\n");
        }
    }
    return RB_OK;
}

#define RB_TC_DSUBS(X) X(2) X(4) X(6) X(8) X(10) X(12) X(16) X(20) X(24) X(30) X(32)

bool dsub_instantiated(size_t dsub)
{
    switch (dsub) {
#define X(D) case D:
        RB_TC_DSUBS(X)
#undef X
        return true;
    default: return false;
    }
}

}  // namespace

}  // namespace rb

// Host-only view of the work plan of the tensor encode kernel, for the CPU tests of its invariants (every column
// group owns at least one CTA, CTA ranges partition [0, ctas), shared memory fits, TMA box constraints).
// out = {gm, n_groups, a_stages, pitch_f, ctas, smem_bytes}; cta_start receives n_groups + 1 entries.  Returns 0 when
// the shape has no plan (the exact kernel is used then).
extern "C" int rb_debug_tensor_plan(size_t M, size_t dsub, size_t n_tiles, int sms, long long *out, unsigned short *cta_start)
{
    const rb::Plan p = rb::make_plan(M, dsub, n_tiles, sms > rb::kMaxGroups ? rb::kMaxGroups : sms);
    if (p.gm == 0) return 0;
    out[0] = p.gm;
    out[1] = p.n_groups;
    out[2] = p.a_stages;
    out[3] = p.pitch_f;
    out[4] = p.ctas;
    out[5] = (long long)p.smem;
    for (int g = 0; g <= p.n_groups; g++) cta_start[g] = p.cta_start[g];
    return 1;
}

namespace rb {

bool tensor_path_supported(const DeviceCodebook &cb)
{
    // k < 256 runs padded to 256 columns; below ~64 centroids the exact kernel (cost proportional to k) is faster
    if (cb.k > (size_t)kCent || cb.k <= 64 || !dsub_instantiated(cb.dsub)) return false;
    return make_plan(cb.M, cb.dsub, 1u << 20, kMaxGroups).gm > 0;
}

bool tensor_call_supported(const DeviceCodebook &cb, const float *x, size_t n, ptrdiff_t ldx)
{
    if (!tensor_path_supported(cb)) return false;
    if (n == 0 || n >= ((size_t)1 << 32)) return false;
    if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) return false;
    if (ldx < (ptrdiff_t)(cb.M * cb.dsub) || (ldx % 4) != 0) return false;
    return true;
}

rb_status TensorOperands::prepare(const DeviceCodebook &cb, cudaStream_t stream)
{
    if (!tensor_path_supported(cb)) return RB_OK;  // stays !ready(): callers use the exact kernel
    kpad = kpad_of((int)cb.dsub);
    const size_t b_bytes = cb.M * (size_t)(kpad / 8) * kCent * 16;
    const size_t c_bytes = (cb.M + 4) * sizeof(float);
    if (!b_tiles) {
        RB_CUDA_TRY(pool_malloc((void **)&b_tiles, b_bytes + c_bytes, stream));
        bytes = b_bytes + c_bytes;
    }
    consts = reinterpret_cast<float *>(reinterpret_cast<char *>(b_tiles) + b_bytes);
    RB_CUDA_TRY(cudaMemsetAsync(consts, 0, c_bytes, stream));
    tc_absmax_kernel<<<148, 256, 0, stream>>>(cb.quantizers, cb.M * cb.k * cb.dsub, cb.cs, cb.M, cb.k, consts);
    RB_LAUNCH_CHECK();
    const unsigned blocks = (unsigned)ceil_div(cb.M * (size_t)kCent, 128);
    switch (cb.dsub) {
#define X(D)                                                                                                         \
    case D:                                                                                                          \
        tc_prepare_kernel<D><<<blocks, 128, 0, stream>>>(cb.quantizers, cb.cs, (int)cb.M, (int)cb.k,                 \
                                                         reinterpret_cast<__half *>(b_tiles), consts);               \
        break;
        RB_TC_DSUBS(X)
#undef X
    default: break;
    }
    RB_LAUNCH_CHECK();
    return RB_OK;
}

void TensorOperands::release()
{
    if (b_tiles) cudaFree(b_tiles);
    b_tiles = nullptr;
    consts = nullptr;
}

void TensorOperands::release_async(cudaStream_t stream)
{
    if (b_tiles) cudaFreeAsync(b_tiles, stream);
    b_tiles = nullptr;
    consts = nullptr;
}

rb_status launch_encode_tensor(const DeviceCodebook &cb, const TensorOperands &tc, const float *x, size_t n, ptrdiff_t ldx,
                               void *codes, int code_width, ptrdiff_t crs, ptrdiff_t ccs, cudaStream_t stream,
                               const RotatedInput *rot)
{
    if (n == 0) return RB_OK;
    if (!tc.ready() || !tensor_call_supported(cb, x, n, ldx)) {
        set_error("tensor encode path does not cover this call (k=%zu, dsub=%zu, ldx=%td)", cb.k, cb.dsub, ldx);
        return RB_ERR_UNSUPPORTED;
    }
    // list of (row, subquantizer) pairs the tensor pass could not decide; when it overflows the whole batch is
    // re-encoded by the exact kernel (gated on the device, no host round trip)
    const size_t units = n * cb.M;
    uint32_t *work = nullptr;
    rb_status st = RB_OK;
    if (rot) {
        // A rotated input cannot fall back to the exact kernel on the same (approximate) buffer.  The flagged rows
        // are collected per subquantizer, [M] counters + [M][n] row indices, and re-decided by a kernel that first
        // re-rotates the subvector exactly (encode_exact.cu).
        if (n > 0xffffffffull || !rotated_recheck_supported(cb, rot->d)) {
            set_error("tensor encode of a rotated batch does not cover this shape (n=%zu, d=%zu)", n, rot->d);
            return RB_ERR_UNSUPPORTED;
        }
        const size_t counters = (cb.M + 3) / 4 * 4;
        RB_CUDA_TRY(pool_malloc((void **)&work, (counters + units) * sizeof(uint32_t), stream));
        uint32_t *counts = work, *rows = work + counters;
        auto body = [&]() -> rb_status {
            RB_CUDA_TRY(cudaMemsetAsync(counts, 0, counters * sizeof(uint32_t), stream));
            switch (cb.dsub) {
#define X(D)                                                                                                         \
    case D:                                                                                                          \
        RB_TRY(launch_t<D>(cb, tc, x, n, ldx, codes, code_width, crs, ccs, rows, counts, 0u, rot, stream));           \
        break;
                RB_TC_DSUBS(X)
#undef X
            default: break;
            }
            if (getenv("RB_TC_STATS")) {
                std::vector<uint32_t> h(cb.M);
                RB_CUDA_TRY(cudaMemcpyAsync(h.data(), counts, cb.M * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
                RB_CUDA_TRY(cudaStreamSynchronize(stream));
                size_t tot = 0;
                for (uint32_t c : h) tot += c;
                fprintf(stderr, "[rb tc] rotated n=%zu M=%zu dsub=%zu: %zu of %zu pairs re-rotated and re-decided exactly (%.4f%%)\n",
                        n, cb.M, cb.dsub, tot, units, 100.0 * tot / (double)units);
            }
            return launch_rotated_recheck(cb, counts, rows, n, rot->x0, rot->ldx0, rot->r, rot->d, codes, code_width, crs, ccs,
                                          stream);
        };
        st = body();
        cudaFreeAsync(work, stream);
        return st;
    }
    size_t cap = units / 8 + 4096;
    if (cap > 0x7fffffffull) cap = 0x7fffffffull;
    RB_CUDA_TRY(pool_malloc((void **)&work, (2 * cap + 4) * sizeof(uint32_t), stream));
    uint32_t *n_pairs = work, *pairs = work + 4;
    auto body = [&]() -> rb_status {
        RB_CUDA_TRY(cudaMemsetAsync(n_pairs, 0, 4 * sizeof(uint32_t), stream));
        switch (cb.dsub) {
#define X(D)                                                                                                         \
    case D:                                                                                                          \
        RB_TRY(launch_t<D>(cb, tc, x, n, ldx, codes, code_width, crs, ccs, pairs, n_pairs, (uint32_t)cap, nullptr,   \
                           stream));                                                                                 \
        break;
            RB_TC_DSUBS(X)
#undef X
        default: break;
        }
        if (getenv("RB_TC_STATS")) {  // debugging aid: how many (row, subquantizer) pairs the tensor pass left undecided
            uint32_t np_host = 0;
            RB_CUDA_TRY(cudaMemcpyAsync(&np_host, n_pairs, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
            RB_CUDA_TRY(cudaStreamSynchronize(stream));
            fprintf(stderr, "[rb tc] n=%zu M=%zu dsub=%zu: %u of %zu pairs re-decided exactly (%.4f%%)\n", n, cb.M, cb.dsub,
                    np_host, units, 100.0 * np_host / (double)units);
        }
        RB_TRY(launch_encode_recheck(cb, x, ldx, pairs, n_pairs, (uint32_t)cap, codes, code_width, crs, ccs, stream));
        RB_TRY(launch_encode_exact_gated(cb, x, n, ldx, codes, code_width, crs, ccs, 0, n_pairs, (uint32_t)cap, stream));
        return RB_OK;
    };
    st = body();
    cudaFreeAsync(work, stream);
    return st;
}

}  // namespace rb
