// encode_tc.cu — nearest-centroid encode on the 5th-generation tensor cores (tcgen05, sm_100a), bit-exact.
//
// Replaces the same reference chain as encode_exact.cu
//   primitives::quantize_batch_into (src/pq/primitives.rs:64-104) -> kmeans::cluster_assignments
//   (src/kmeans.rs:133-159) -> SquaredEuclideanDistance<Ix2> (src/linalg.rs:150-180)
// for k <= 256 centroids.  The reference's decision for one (row, subquantizer) is
//   argmin_j  d_j,   d_j = fl(fl(xs + cs_j) - fl(2 dp_j)),  dp_j a sequential FP32 FMA chain      (linalg.rs:167-176)
// which cannot be reproduced bit for bit by a tensor-core contraction.  The tensor pass is therefore a FILTER: it
// proves, per (row, subquantizer), which centroids cannot be the reference's argmin, and everything it cannot rule
// out is decided by the reference's exact expression tree.
//
//   1. Scores  s~_j = [xs +] cs_j - 2 x.c_j  for all 256 centroids come from tcgen05.mma kind::f16 (M=128 rows,
//      N=256) with F16 ACCUMULATORS in tensor memory.  Operands are two-limb FP16 splits of the inputs scaled by
//      a power of two PER SUBQUANTIZER (x = xh + xl, c = ch + cl); the products xl.ch, xh.cl, xh.ch, the two-limb
//      ||c||^2 and (when a K column is free) ||x||^2 are laid along K with the small products FIRST, so that the
//      only roundings of magnitude are the F16 roundings of the last instruction(s): the score is exact to about
//      2^-11 of ITS OWN magnitude (measured: the final conversion is round-to-nearest, scripts/microbench/probe2.cu).
//   2. Each epilogue thread owns one row (one TMEM lane), reads the 256 scores packed two per register
//      (tcgen05.ld.pack::16b) and reduces them with packed 3-input FP16 minima (VHMNMX: 4 scores per instruction,
//      half the ALU-pipe work of an FP32 scan) along two partitions of the index set — 16 chains (j mod 16) and
//      32 half-blocks (16 consecutive centroids, even / odd).
//   3. thr = m1 + 2^-9 |m1| + margin bounds the tensor score of the reference's argmin from above (m1 = smallest
//      score; margin from ||x||^2 and max_j ||c_j||^2, see margin_coef).  The chain / block minima at or below thr
//      give two flag words; every centroid outside (flagged chains) x (flagged blocks) is PROVEN not to be the argmin.
//      One candidate (~99 % of the pairs on Gaussian data): its index is the code.  Otherwise the pair goes to a
//      list and launch_encode_candidates evaluates just the candidates with the reference's exact tree
//      (non-finite or out-of-range values: all 256).
//   The emitted codes are therefore identical to the exact kernel's (and the oracle's) for every input.
//
// Data flow of one CTA (persistent, one per SM, 16 warps = four per SM sub-partition, 128 registers each):
//   warps 12-15 converters: thread = row; split the FP32 subvector into FP16 limbs, write the A operand in the
//               no-swizzle K-major core-matrix layout, publish the row's margin.  Warp 12 is also the producer: the
//               group's B operands once (cp.async.bulk), then one TMA tensor-map tile load per row tile (128 rows
//               x the group's column slice, pitch an odd multiple of 16 bytes)
//   warps 0-11  epilogue, three sets of four warps taking the units in turn (unit u: set u mod 3, accumulator
//               u mod 2); a set holds an accumulator only while it scans, so its certificate overlaps the other
//               sets' scans.  There is no separate MMA warp: when a set has finished scanning unit u, ONE of its
//               warps waits for the other three, then issues the tcgen05.mma (K/16 instructions) and commits of
//               unit u + 2 into the accumulator it has just released (a dedicated issuer warp slowed the epilogue
//               warps of its sub-partition by a third and added a wake-up to every accumulator hand-off)
// A rotated input (x ~ x0 . R from project_tc.cu, template parameter ROT) widens the margin by the rotation's error
// bound and collects the undecided rows per subquantizer for an exact re-rotation (encode_tc.cuh RotatedInput).
// Bounds (C2: 2M x 300, M=30): HBM 4*d + M bytes per vector is the roofline (0.38 ms); see DESIGN.md.
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
constexpr int kMaxCandidates = 16;  // (chains below thr) x (blocks below thr) a candidate-list entry may name
constexpr int kSets = 3;  // epilogue sets (four warps each, TMEM lane quarter = warp % 4)
constexpr int kThreads = 32 * (4 * kSets + 4);
// Warp roles: 12 epilogue warps (TMEM lane quarter = warp % 4; warp 5 * s of set s also issues MMAs), 4 converter
// warps (the first also drives the TMA loads).
constexpr int kWarpConv0 = 4 * kSets;
constexpr int kSmemLimit = 227 * 1024;

// K layout of the augmented operands for subvector width D (all offsets in halves, all even):
//   [0, D)            A: xl      B: -2 ch     small products
//   [D, 2D)           A: xh      B: -2 cl     small products
//   zeros
//   [main, main + D)  A: xh      B: -2 ch     the main products, right-aligned at KPAD together with
//   [.., +2)          A: 1 1     B: csh csl   ||c||^2 in two limbs and, when two more columns are free without
//   [.., +2)          A: xsh xsl B: 1 1       costing another instruction, ||x||^2 (makes the score the distance)
// F16 accumulators round after every K = 16 instruction; with this order the instructions before the last
// n_main(D) only ever hold sums of small products (|.| <= 2^-9 |x||c|), whose F16 rounding is negligible.
__host__ __device__ constexpr int kpad_of(int dsub) { return ((3 * dsub + 2 + 15) / 16) * 16; }
__host__ __device__ constexpr int nxs_of(int dsub)
{
    return (3 * dsub + 4 <= kpad_of(dsub) && (dsub + 4 + 15) / 16 == (dsub + 2 + 15) / 16) ? 2 : 0;
}
__host__ __device__ constexpr int main_off(int dsub) { return kpad_of(dsub) - (dsub + 2 + nxs_of(dsub)); }
__host__ __device__ constexpr int n_main_of(int dsub) { return (dsub + 2 + nxs_of(dsub) + 15) / 16; }

// Error budget of one tensor score against the reference's d_j (minus the common ||x||^2 when it is not in K), in
// units of (XS + CS_j), scaled inputs X = S x, C = S c (S a power of two):
//   limb split of the products  3 * 2^-22      (|X_t C_t - xl ch - xh cl - xh ch| <= 3 * 2^-22 |X_t C_t|, times 2,
//                                               Cauchy-Schwarz, |X||C| <= (XS + CS) / 2)
//   ||c||^2 limbs               2^-22
//   tensor-core accumulation    2^-20          (measured exact to the final rounding for K = 16; generous)
//   small-product instructions  <= 6 * 2^-21   (F16 rounding of a partial sum below 2^-9 (XS + CS))
//   FP16 subnormal limbs        <= 1.5 * 2^-20 (absolute 2^-24 sqrt(dsub (XS + CS)); CS_max >= 16 in scaled units)
//   the reference's own FP32 roundings (two adds, the FMA chain, the norms)  <= (4 + dsub) 2^-24 <= 2.25 * 2^-20
// together below 8 * 2^-20; the bound is applied to both the winner and the argmin: 2^-16 covers twice that.  Every
// instruction of the main block except the last adds an F16 rounding of a partial sum of magnitude <= 2 (XS + CS):
// 2^-10 each (again for two scores: 2.05 * 2^-10).  The LAST rounding is relative to the score itself and is
// the 2^-9 |m1| term of the threshold (see the epilogue).
__host__ __device__ constexpr float margin_coef(int dsub)
{
    return 1.52587890625e-5f + (float)(n_main_of(dsub) - 1) * 2.05f * 9.765625e-4f;
}

// ---------------------------------------------------------------------------------------------------------
// operand preparation (runs when a codebook is created / after every k-means update)
// ---------------------------------------------------------------------------------------------------------
// consts: [M][4] = { max_j ||c_j||^2 * scale^2, scale, scale^2, bad } per subquantizer.  One block per subquantizer.
__global__ void __launch_bounds__(256) tc_scale_kernel(const float *__restrict__ q, const float *__restrict__ cs, int k, int dsub,
                                                       float *__restrict__ consts)
{
    const int m = blockIdx.x;
    __shared__ unsigned s_amax, s_csmax;
    if (threadIdx.x == 0) s_amax = s_csmax = 0u;
    __syncthreads();
    unsigned amax = 0, csmax = 0;
    const float *qm = q + (size_t)m * k * dsub;
    for (int i = threadIdx.x; i < k * dsub; i += blockDim.x) amax = max(amax, __float_as_uint(qm[i]) & 0x7fffffffu);
    for (int i = threadIdx.x; i < k; i += blockDim.x) csmax = max(csmax, __float_as_uint(cs[(size_t)m * k + i]) & 0x7fffffffu);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        amax = max(amax, __shfl_xor_sync(0xffffffffu, amax, off));  // NaN / Inf compare above every finite value
        csmax = max(csmax, __shfl_xor_sync(0xffffffffu, csmax, off));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&s_amax, amax);
        atomicMax(&s_csmax, csmax);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const float a = __uint_as_float(s_amax), c = __uint_as_float(s_csmax);
        const bool bad = !(a < 3.0e38f) || !(c < 3.0e38f);  // NaN or Inf somewhere in this codebook
        float scale = 1.f;
        if (!bad && a != 0.f) {
            int e = 3 - (ilogbf(a) + 1);  // a * 2^e in [4, 8)
            e = max(-60, min(60, e));
            scale = scalbnf(1.f, e);
        }
        const float scale2 = scale * scale;
        consts[4 * m + 0] = c * scale2;
        consts[4 * m + 1] = scale;
        consts[4 * m + 2] = scale2;
        consts[4 * m + 3] = (bad || !(c * scale2 < 3.0e38f)) ? 1.f : 0.f;
    }
}

template <int DSUB>
__global__ void tc_prepare_kernel(const float *__restrict__ q, const float *__restrict__ cs, int M, int k,
                                  __half *__restrict__ bop, const float *__restrict__ consts)
{
    constexpr int KPAD = kpad_of(DSUB), NCH = KPAD / 8, MAIN = main_off(DSUB), NXS = nxs_of(DSUB);
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (m, j)
    if (idx >= M * kCent) return;
    const int m = idx / kCent, j = idx % kCent;
    const float scale = consts[4 * m + 1], scale2 = consts[4 * m + 2];
    __half kv[KPAD];
#pragma unroll
    for (int t = 0; t < KPAD; t++) kv[t] = __float2half_rn(0.f);
    if (j >= k) {
        // k < 256: the unused columns get a zero centroid with the constant score kPadScore, far above any real
        // score of a row the tensor pass is allowed to decide (rows with large norms are decided exactly)
        kv[MAIN + DSUB] = __float2half_rn(kPadScore);
    } else {
        const float *c = q + ((size_t)m * k + j) * DSUB;
#pragma unroll
        for (int t = 0; t < DSUB; t++) {
            const float cv = c[t] * scale;
            const __half h = __float2half_rn(cv);
            const __half l = __float2half_rn(cv - __half2float(h));
            const __half h2 = __float2half_rn(-2.f * __half2float(h));  // exact (|cv| < 8)
            const __half l2 = __float2half_rn(-2.f * __half2float(l));
            kv[t] = h2;          // x  xl
            kv[DSUB + t] = l2;   // x  xh
            kv[MAIN + t] = h2;   // x  xh
        }
        const float csv = cs[(size_t)m * k + j] * scale2;
        const __half ch = __float2half_rn(csv);
        kv[MAIN + DSUB] = ch;
        kv[MAIN + DSUB + 1] = __float2half_rn(csv - __half2float(ch));
        if (NXS) kv[MAIN + DSUB + 2] = kv[MAIN + DSUB + 3] = __float2half_rn(1.f);
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
    const float *consts;  // [M][4]: csmax * scale^2 | scale | scale^2 | bad, per subquantizer
    void *codes;
    int code_width;
    long long crs, ccs;
    // Undecided (row, subquantizer) pairs, 4 words each: row, m, chain flags, block flags (bit layout in the epilogue;
    // all ones = every centroid is a candidate).  Every epilogue warp appends to its own region of `region_cap`
    // entries (region = 12 * CTA + warp), so no atomics are involved: region_counts[region] receives the final count,
    // *overflow is set when a region filled up (the gated exact kernel then redoes the batch).
    uint32_t *cands, *region_counts, *overflow;
    uint32_t region_cap;
    int M, k, gm, n_groups, a_stages, pitch_f;
    float xs_limit;  // rows with ||x * scale||^2 at or above this are decided exactly
    // x is an APPROXIMATE rotation of other rows (project_tc.cu): every component of a row may be off by
    // rowerr[row] + perr_floor / *perr_sx, which widens the margin; rowerr == nullptr: x is exact
    uint32_t *bucket_counts, *bucket_rows;  // rotated input: undecided (row, candidate flags) per subquantizer, [M] and [M][n][2]
    const float *rowerr;
    const float *perr_sx;  // device scalar: the rotation's operand scale
    float perr_floor;
    long long n_tiles;
    long long *trace;  // debugging aid (RB_TC_TRACE): per-role clock64 stamps of CTA 0, else nullptr
    unsigned short cta_start[kMaxGroups + 1];  // CTAs [cta_start[g], cta_start[g+1]) own column group g
};

// packed FP16 minimum of two / three registers (two scores each); ptxas fuses the pair into one 3-input VHMNMX
__device__ __forceinline__ uint32_t hmin2(uint32_t a, uint32_t b)
{
    uint32_t r;
    asm("min.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ uint32_t hmin3(uint32_t a, uint32_t b, uint32_t c) { return hmin2(hmin2(a, b), c); }
// minimum of 8 packed registers = 16 consecutive scores, kept as (even minimum, odd minimum)
__device__ __forceinline__ uint32_t hmin8(const uint32_t *v)
{
    return hmin3(hmin3(v[0], v[1], v[2]), hmin3(v[3], v[4], v[5]), hmin2(v[6], v[7]));
}
// packed FP16 compare: 0xffff in each half where a <= b (false for NaN)
__device__ __forceinline__ uint32_t hle2(uint32_t a, uint32_t b)
{
    uint32_t r;
    asm("set.le.u32.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ float h_lo(uint32_t v) { return __half2float(__ushort_as_half((unsigned short)(v & 0xffffu))); }
__device__ __forceinline__ float h_hi(uint32_t v) { return __half2float(__ushort_as_half((unsigned short)(v >> 16))); }

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
    // acc_full has one barrier per (set, accumulator) pair = unit index mod 2 kSets: the units u and u + kSets of one set
    // use DIFFERENT accumulators, so the MMA of u + kSets can complete before the set has even started to wait for u
    // (its certificate of u - kSets may take long) -- on one barrier that is two completions, which a parity wait
    // cannot tell from none: the set would sleep forever (this deadlocked ~1 launch in 50 of the rotated encode).
    // Units u and u + 2 kSets share their accumulator and are therefore ordered.  acc_empty: one per accumulator.
    uint64_t *x_full = bars, *x_empty = bars + 2, *a_full = bars + 4, *a_empty = bars + 8, *acc_full = bars + 12,
             *acc_empty = bars + 12 + 2 * kSets, *b_full = bars + 14 + 2 * kSets;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 15 + 2 * kSets);
    static_assert(16 + 2 * kSets <= 24, "barrier block of the shared-memory plan");

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(&x_full[i], 1);
            mbar_init(&x_empty[i], 4);
            mbar_init(&acc_empty[i], 4);
        }
        for (int i = 0; i < 2 * kSets; i++) mbar_init(&acc_full[i], 1);
        for (int i = 0; i < 4; i++) {
            mbar_init(&a_full[i], 4);
            mbar_init(&a_empty[i], 1);
        }
        mbar_init(b_full, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
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
    // units of this CTA: (tile, subquantizer) pairs in the order the converters produce them
    const uint32_t n_units = (uint32_t)(t_first < p.n_tiles ? ((p.n_tiles - t_first + t_stride - 1) / t_stride) * gm_cur : 0);

    if (warp >= kWarpConv0) {
        // ===================== converters (thread = row); warp kWarpConv0 also drives the TMA loads =====================
        constexpr int MAIN = main_off(DSUB), NXS = nxs_of(DSUB);
        const int row = (warp - kWarpConv0) * 32 + lane;
        const float4 *consts4 = reinterpret_cast<const float4 *>(p.consts) + g * p.gm;
        uint32_t u = 0, as = 0, aph = 0, li = 0;
        const bool producer = warp == kWarpConv0 && lane == 0;
        if (producer) {  // B operands of the group and the first two tiles
            prefetch_tensormap(&tmap);
            mbar_arrive_expect_tx(b_full, (uint32_t)gm_cur * B_BYTES);
            for (int ml = 0; ml < gm_cur; ml++)
                bulk_g2s(sB + (size_t)ml * B_BYTES, p.bop + (size_t)(g * p.gm + ml) * (B_BYTES / 2), B_BYTES, b_full);
            long long t = t_first;
            for (int st = 0; st < kXStages && t < p.n_tiles; st++, t += t_stride) {
                // one TMA tile load: box = 128 rows x pitch floats; rows / columns outside the matrix arrive as zeros
                mbar_arrive_expect_tx(&x_full[st], (uint32_t)xs_bytes);
                tma_load_2d(sX + (size_t)st * xs_bytes, &tmap, g * p.gm * DSUB, (int)(t * kTile), &x_full[st]);
            }
        }
        __syncwarp();
        RB_PH_BEGIN();
        for (long long t = t_first; t < p.n_tiles; t += t_stride, li++) {
            const int stage = (int)(li & 1);
            RB_PH(0);
            mbar_wait(&x_full[stage], (li >> 1) & 1, 1);
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
                    const float4 cm = consts4[ml0 + h];  // csmax * scale^2 | scale | scale^2 | bad  (this subquantizer)
                    const float scale = cm.y;
                    unsigned long long xs2 = 0ull;  // ||x * scale||^2 as two interleaved partial sums
                    uint32_t hw[DSUB / 2], lw[DSUB / 2];
#pragma unroll
                    for (int t2 = 0; t2 < DSUB / 2; t2++) {
                        const float a0 = sub[2 * t2], a1 = sub[2 * t2 + 1];
                        // packed FP32 (mul / fma .f32x2) on the element pair; two-limb FP16 split, two elements per
                        // conversion instruction
                        const unsigned long long s2 = mul2(pack2(a0, a1), pack2(scale, scale));
                        xs2 = fma2(s2, s2, xs2);
                        const float s0 = lo2(s2), s1 = hi2(s2);
                        const __half2 hh = __floats2half2_rn(s0, s1);
                        const float2 hf = __half22float2(hh);
                        const unsigned long long l2 = fma2(pack2(hf.x, hf.y), pack2(-1.f, -1.f), s2);  // s - hf, one rounding
                        const __half2 ll = __floats2half2_rn(lo2(l2), hi2(l2));
                        hw[t2] = *reinterpret_cast<const uint32_t *>(&hh);
                        lw[t2] = *reinterpret_cast<const uint32_t *>(&ll);
                    }
                    // NaN in x makes xs NaN, Inf makes it Inf; |x * scale| <= sqrt(xs) must stay inside the FP16 range,
                    // scores (<= (|x| + |c|)^2 with |c| <= 8 sqrt(dsub)) inside it as well and, for padded codebooks,
                    // far below kPadScore: xs_limit = 10^4 gives scores below 21 100.  Other rows are decided exactly.
                    const float xs_sc = lo2(xs2) + hi2(xs2);
                    const bool bad = (cm.w != 0.f) || !(xs_sc < p.xs_limit);
                    // K layout [xl | xh | 0 ... | xh | 1 1 | xsh xsl] as 32-bit words (two halves each), see kpad_of
                    uint32_t w[KPAD / 2];
#pragma unroll
                    for (int i = 0; i < KPAD / 2; i++) w[i] = 0u;
#pragma unroll
                    for (int i = 0; i < DSUB / 2; i++) {
                        w[i] = lw[i];
                        w[DSUB / 2 + i] = hw[i];
                        w[MAIN / 2 + i] = hw[i];
                    }
                    w[(MAIN + DSUB) / 2] = 0x3c003c00u;  // (1.0, 1.0)
                    if constexpr (NXS != 0) {
                        // ||x||^2 is the same for every centroid: any value works as long as all 256 scores get the
                        // same one; two limbs keep the score a small number near the winner
                        const __half xh = __float2half_rn(xs_sc);
                        const __half xl = __float2half_rn(xs_sc - __half2float(xh));
                        w[(MAIN + DSUB) / 2 + 1] = (uint32_t)__half_as_ushort(xh) | ((uint32_t)__half_as_ushort(xl) << 16);
                    }
                    float marg = margin_coef(DSUB) * (xs_sc + cm.x);
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
                        margc = (NXS != 0 ? 0.f : xs_sc) + eb;  // the score already holds ||x||^2 when it is in K
                    }
                    if (bad || !(marg < 3.0e38f) || !(margb < 3.0e38f))
                        marg = __int_as_float(0x7fc00000);  // NaN: always re-decide exactly
                    RB_PH(2);
                    mbar_wait(&a_empty[as], aph ^ 1, 2);
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
            if (producer && t + 2 * t_stride < p.n_tiles) {
                // refill this stage with the tile after next once all four converter warps have left it
                mbar_wait(&x_empty[stage], (li >> 1) & 1, 3);
                mbar_arrive_expect_tx(&x_full[stage], (uint32_t)xs_bytes);
                tma_load_2d(sX + (size_t)stage * xs_bytes, &tmap, g * p.gm * DSUB, (int)((t + 2 * t_stride) * kTile), &x_full[stage]);
            }
            __syncwarp();
        }
        RB_PH_END();
    } else {
        // ===================== epilogue (thread = row = TMEM lane) =====================
        const int set = warp >> 2;
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16);
        // this set's units are u = set, set + kSets, ...: (tile, ml) advance by kSets subquantizers with wrap-around
        long long t = t_first;
        int ml = set;
        while (ml >= gm_cur && t < p.n_tiles) {
            ml -= gm_cur;
            t += t_stride;
        }
        uint32_t n_flagged = 0;  // entries this warp has appended to its region (warp-uniform)
        uint32_t full_ix = (uint32_t)set, full_ph = 0;  // acc_full barrier (u mod 2 kSets) and phase ((u / 2 kSets) & 1) of unit u
        // MMA issue: the whole warp runs the code and one elected lane issues, so descriptors, phases and barrier
        // addresses stay warp-uniform.  The shared-memory descriptors differ only in the 14-bit start-address field of
        // their low word.  Unit v uses A stage v mod S, accumulator v mod 2, and completes on acc_full[v mod kSets].
        const bool issuer = q == set;
        const uint32_t idesc = idesc_f16(kTile, kCent, 0) & ~(3u << 4);  // D format field = 0: F16 accumulator
        const uint64_t a_desc0 = smem_desc_kmajor(smem_u32(sA), kTile * 16, 128);
        const uint64_t b_desc0 = smem_desc_kmajor(smem_u32(sB), kCent * 16, 128);
        constexpr uint32_t A_STEP = A_BYTES >> 4, B_STEP = B_BYTES >> 4;                // per stage / per subquantizer
        constexpr uint32_t A_KS = (2 * kTile * 16) >> 4, B_KS = (2 * kCent * 16) >> 4;  // per K = 16 slice
        // (unit, its A stage and phase, subquantizer, acc_full barrier) of the next unit this warp issues
        uint32_t iv = 0, i_as = 0, i_aph = 0, i_ml = 0, i_set = 0;
        auto issue_at = [&](uint32_t v) {  // position the issue state on unit v
            iv = v;
            i_as = v % (uint32_t)S;
            i_aph = (v / (uint32_t)S) & 1u;
            i_ml = v % (uint32_t)gm_cur;
            i_set = v % (uint32_t)(2 * kSets);
        };
        auto issue_unit = [&]() {  // the accumulator iv mod 2 is free
            mbar_wait(&a_full[i_as], i_aph, 4);
            tc_fence_after();
            const uint32_t a_lo = (uint32_t)a_desc0 + i_as * A_STEP, b_lo = (uint32_t)b_desc0 + i_ml * B_STEP;
#pragma unroll
            for (int ks = 0; ks < KPAD / 16; ks++)
                mma_f16_ss_lohi_warp(tmem_base + (iv & 1) * kCent, a_lo + ks * A_KS, (uint32_t)(a_desc0 >> 32), b_lo + ks * B_KS,
                                     (uint32_t)(b_desc0 >> 32), idesc, ks > 0 ? 1u : 0u);
            tc_commit_warp(&acc_full[i_set]);
            tc_commit_warp(&a_empty[i_as]);
        };
        if (issuer) {
            mbar_wait(b_full, 0, 5);
            if (set >= 1) {  // units 0 and 1 start the pipeline (sets 1 and 2 issue them before their first scan)
                issue_at((uint32_t)set - 1u);
                if (iv < n_units) issue_unit();
            }
            issue_at((uint32_t)set + 2u);  // from now on: after scanning own unit u (= set, set + kSets, ...) issue u + 2
        }
        RB_PH_BEGIN();
        for (uint32_t u = (uint32_t)set; t < p.n_tiles; u += kSets) {
            const long long grow = t * kTile + row;
            const int m = g * p.gm + ml;
            {
                const uint32_t buf = u & 1;
                const uint32_t taddr = taddr0 + buf * kCent;
                RB_PH(0);
                // the previous unit may have ended in a divergent branch (flagged rows): the tcgen05.ld / elect.sync
                // instructions below are warp-collective and need the warp converged
                __syncwarp();
                mbar_wait(&acc_full[full_ix], full_ph, 6);
                __syncwarp();
                RB_PH(1);
                tc_fence_after();
                // 256 F16 scores, two per register: register r of a 64-column load = columns (2r, 2r + 1).
                //   A[i]  (i < 8)  = minima over registers r = i mod 8    -> 16 chains  j mod 16 = (2i, 2i + 1)
                //   B[b]  (b < 16) = minima over registers 8 (b mod 4) .. + 8 of load b / 4 -> block b = j / 16,
                //                    kept as (minimum over even j, minimum over odd j)
                uint32_t A[8], B[16];
                uint32_t v0[32], v1[32];
                tmem_ld32_pack16(taddr, v0);
                tmem_wait_ld(v0);
                tmem_ld32_pack16(taddr + 64, v1);
#pragma unroll
                for (int gq = 0; gq < 4; gq++) B[gq] = hmin8(v0 + 8 * gq);
#pragma unroll
                for (int a = 0; a < 8; a++) A[a] = hmin3(hmin2(v0[a], v0[8 + a]), v0[16 + a], v0[24 + a]);
                tmem_wait_ld(v1);
                tmem_ld32_pack16(taddr + 128, v0);
#pragma unroll
                for (int gq = 0; gq < 4; gq++) B[4 + gq] = hmin8(v1 + 8 * gq);
#pragma unroll
                for (int a = 0; a < 8; a++) A[a] = hmin3(hmin3(A[a], v1[a], v1[8 + a]), v1[16 + a], v1[24 + a]);
                tmem_wait_ld(v0);
                tmem_ld32_pack16(taddr + 192, v1);
#pragma unroll
                for (int gq = 0; gq < 4; gq++) B[8 + gq] = hmin8(v0 + 8 * gq);
#pragma unroll
                for (int a = 0; a < 8; a++) A[a] = hmin3(hmin3(A[a], v0[a], v0[8 + a]), v0[16 + a], v0[24 + a]);
                tmem_wait_ld(v1);
                // the accumulator is free again
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[buf]);
                if (issuer && iv < n_units) {
                    // unit u + 2 goes into the accumulator this set has just read: wait for the set's other warps
                    mbar_wait(&acc_empty[buf], (u >> 1) & 1, 7);
                    issue_unit();
                    // advance by kSets units without divisions
                    iv += (uint32_t)kSets;
                    i_set = i_set >= (uint32_t)kSets ? i_set - (uint32_t)kSets : i_set + (uint32_t)kSets;
                    i_as += (uint32_t)kSets;
                    while (i_as >= (uint32_t)S) {
                        i_as -= (uint32_t)S;
                        i_aph ^= 1;
                    }
                    i_ml += (uint32_t)kSets;
                    while (i_ml >= (uint32_t)gm_cur) i_ml -= (uint32_t)gm_cur;
                }
#pragma unroll
                for (int gq = 0; gq < 4; gq++) B[12 + gq] = hmin8(v1 + 8 * gq);
#pragma unroll
                for (int a = 0; a < 8; a++) A[a] = hmin3(hmin3(A[a], v1[a], v1[8 + a]), v1[16 + a], v1[24 + a]);
                RB_PH(2);
                float marg = sMarg[(u % kMargRing) * kTile + row];
                float margb = 0.f, margc = 0.f;
                if constexpr (ROT) {
                    margb = sMargB[(u % kMargRing) * kTile + row];
                    margc = sMargC[(u % kMargRing) * kTile + row];
                }

                // smallest score: every score is in exactly one chain
                const uint32_t mp = hmin3(hmin3(A[0], A[1], A[2]), hmin3(A[3], A[4], A[5]), hmin2(A[6], A[7]));
                const float m1 = fminf(h_lo(mp), h_hi(mp));
                // Threshold.  With e the bound on |score - reference value| before the last F16 rounding (margin = 2 e,
                // margin_coef) and the last rounding at most 2^-11 of the score itself: the reference's argmin j*
                // satisfies  s~(j*) <= m1 + 2 e + 2^-11 (|m1| + |s~(j*)|) (1 + 2^-10)  <  m1 + 2^-9 |m1| + margin.
                // Every centroid whose score is above thr (rounded UP to FP16) is therefore not the reference's argmin.
                if constexpr (ROT) marg = fmaf(margb, sqrtf(fmaxf(margc + m1, 0.f)), marg);
                const float thr = fmaf(fabsf(m1), 1.953125e-3f, m1 + marg);
                const uint32_t thr1 = (uint32_t)__half_as_ushort(__float2half_ru(thr));
                const uint32_t thr2 = thr1 | (thr1 << 16);
                // Which chain / block minima are <= thr: one packed compare (0xffff per half that passes; NaN never
                // passes) and one logic op per register collect the flags in two words,
                //   ma: bit i = chain 2 i, bit 16 + i = chain 2 i + 1;   mb: bit b = block b over even j, bit 16 + b = odd j
                uint32_t ma4[4] = {0u, 0u, 0u, 0u}, mb4[4] = {0u, 0u, 0u, 0u};  // four partial words: short dependency chains
#pragma unroll
                for (int i = 0; i < 8; i++) ma4[i & 3] |= hle2(A[i], thr2) & (0x00010001u << i);
#pragma unroll
                for (int b = 0; b < 16; b++) mb4[b & 3] |= hle2(B[b], thr2) & (0x00010001u << b);
                const uint32_t ma = (ma4[0] | ma4[1]) | (ma4[2] | ma4[3]), mb = (mb4[0] | mb4[1]) | (mb4[2] | mb4[3]);
                // exactly one chain and exactly one half-block at or below thr: that centroid is the reference's argmin
                const bool finite = fabsf(m1) < 6.0e4f && thr < 6.0e4f;  // false for NaN margins / scores as well
                bool certain = finite && ma != 0u && mb != 0u && (ma & (ma - 1u)) == 0u && (mb & (mb - 1u)) == 0u;
                const int pa = 31 - __clz((int)ma), pb = 31 - __clz((int)mb);
                unsigned code = (unsigned)(((pb & 15) << 4) | ((pa & 15) << 1) | (pa >> 4));
                certain = certain && code < (unsigned)p.k;  // padded codebooks: never a padding column
                const bool valid = grow < p.n;
                if (valid) store_code(p.codes, p.code_width, grow * p.crs + (long long)m * p.ccs, certain ? code : 0u);
                const bool flagged = !certain && valid;
                {
                    // append to this warp's private region: positions from a ballot, the count lives in a register
                    const unsigned fl = __ballot_sync(0xffffffffu, flagged);
                    if (fl != 0u) {
                        const bool few = finite && ma != 0u && mb != 0u &&
                                         __popc(ma) * __popc((mb | (mb >> 16)) & 0xffffu) <= kMaxCandidates;
                        const uint32_t pos = n_flagged + __popc(fl & ((1u << lane) - 1u));
                        if (flagged) {
                            if (pos < p.region_cap) {
                                uint4 *e = reinterpret_cast<uint4 *>(p.cands) + (size_t)(blockIdx.x * (4 * kSets) + warp) * p.region_cap + pos;
                                *e = make_uint4((uint32_t)grow, (uint32_t)m, few ? ma : 0xffffffffu, few ? mb : 0xffffffffu);
                            } else if constexpr (ROT) {
                                // a rotated batch has no exact kernel to fall back on: straight to the re-rotation
                                // buckets (a row is flagged at most once per subquantizer: slot < n)
                                const uint32_t slot = atomicAdd(&p.bucket_counts[m], 1u);
                                reinterpret_cast<uint2 *>(p.bucket_rows)[(size_t)m * (size_t)p.n + slot] =
                                    make_uint2((uint32_t)grow, 0xffffffffu);  // (row, every centroid is a candidate)
                            } else {
                                *p.overflow = 1u;
                            }
                        }
                        n_flagged += __popc(fl);
                    }
                }
                RB_PH(3);
            }
            ml += kSets;
            while (ml >= gm_cur && t < p.n_tiles) {
                ml -= gm_cur;
                t += t_stride;
            }
            if (full_ix >= (uint32_t)kSets) {  // u + kSets: the other barrier of this set; every second step a new phase
                full_ix -= (uint32_t)kSets;
                full_ph ^= 1;
            } else {
                full_ix += (uint32_t)kSets;
            }
        }
        if (lane == 0) p.region_counts[blockIdx.x * (4 * kSets) + warp] = min(n_flagged, p.region_cap);
        RB_PH_END();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
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

// where the undecided pairs go: plain input -> per-warp regions of one list; rotated input -> per-subquantizer buckets
struct Undecided {
    uint32_t *cands = nullptr, *region_counts = nullptr, *overflow = nullptr;
    uint32_t region_cap = 0;
    uint32_t *bucket_counts = nullptr, *bucket_rows = nullptr;
};

// pairs one epilogue warp can meet: its share of the busiest CTA's (tile, subquantizer) units, 32 rows each
size_t pairs_per_epilogue_warp(const Plan &plan, size_t M, size_t n_tiles)
{
    size_t most = 0;
    for (int g = 0; g < plan.n_groups; g++) {
        const size_t ctas = plan.cta_start[g + 1] - plan.cta_start[g];
        const size_t width = g + 1 < plan.n_groups ? (size_t)plan.gm : M - (size_t)(plan.n_groups - 1) * plan.gm;
        const size_t units = ceil_div(n_tiles, ctas) * width;
        most = units > most ? units : most;
    }
    return ceil_div(most, (size_t)kSets) * 32;
}

template <int DSUB>
rb_status launch_t(const DeviceCodebook &cb, const TensorOperands &tc, const float *x, size_t n, ptrdiff_t ldx, void *codes,
                   int code_width, ptrdiff_t crs, ptrdiff_t ccs, const Undecided &und, const RotatedInput *rot,
                   cudaStream_t stream)
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
    p.cands = und.cands;
    p.region_counts = und.region_counts;
    p.overflow = und.overflow;
    p.region_cap = und.region_cap;
    p.k = (int)cb.k;
    p.M = (int)cb.M;
    p.gm = plan.gm;
    p.n_groups = plan.n_groups;
    p.a_stages = plan.a_stages;
    p.pitch_f = plan.pitch_f;
    p.n_tiles = (long long)n_tiles;
    p.xs_limit = 1.0e4f;
    p.trace = nullptr;
    p.bucket_counts = und.bucket_counts;
    p.bucket_rows = und.bucket_rows;
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
        fprintf(stderr, "  epilogue (warps 0-11): loop | wait acc_full | scan (+ MMA issue on warps 0, 5, 10) | certificate+store\n");
        fprintf(stderr, "  converter (warps 12-15): loop | wait x_full | load+convert | wait a_empty | write+publish\n");
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
    const size_t c_bytes = cb.M * 4 * sizeof(float);
    if (!b_tiles) {
        RB_CUDA_TRY(pool_malloc((void **)&b_tiles, b_bytes + c_bytes, stream));
        bytes = b_bytes + c_bytes;
    }
    consts = reinterpret_cast<float *>(reinterpret_cast<char *>(b_tiles) + b_bytes);
    tc_scale_kernel<<<(unsigned)cb.M, 256, 0, stream>>>(cb.quantizers, cb.cs, (int)cb.k, (int)cb.dsub, consts);
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

namespace {

// one launch of the tensor pass + its exact follow-up over rows [0, n), n * M < 2^31
rb_status encode_tensor_chunk(const DeviceCodebook &cb, const TensorOperands &tc, const float *x, size_t n, ptrdiff_t ldx,
                              void *codes, int code_width, ptrdiff_t crs, ptrdiff_t ccs, cudaStream_t stream)
{
    // The list of (row, subquantizer) pairs the tensor pass could not decide (about 1 % on Gaussian data, nearly all
    // with two or three candidates): one private region per epilogue warp, sized for an eighth of the pairs the
    // warp can meet.  When a region fills up, the whole chunk is re-encoded by the gated exact kernel (on the
    // device, no host round trip).
    const size_t n_tiles = ceil_div(n, (size_t)kTile);
    const Plan plan = make_plan(cb.M, cb.dsub, n_tiles, device_sm_count(), 1);
    const size_t regions = (size_t)plan.ctas * 4 * kSets;
    const size_t region_cap = pairs_per_epilogue_warp(plan, cb.M, n_tiles) / 8 + 64;
    const size_t head = 4 + (regions + 3) / 4 * 4;  // overflow flag (+ pad) | region counts; entries stay 16-byte aligned
    uint32_t *work = nullptr;
    RB_CUDA_TRY(pool_malloc((void **)&work, (head + 4 * regions * region_cap) * sizeof(uint32_t), stream));
    Undecided und;
    und.overflow = work;
    und.region_counts = work + 4;
    und.cands = work + head;
    und.region_cap = (uint32_t)region_cap;
    auto body = [&]() -> rb_status {
        RB_CUDA_TRY(cudaMemsetAsync(work, 0, head * sizeof(uint32_t), stream));
        switch (cb.dsub) {
#define X(D)                                                                                                         \
    case D:                                                                                                          \
        RB_TRY(launch_t<D>(cb, tc, x, n, ldx, codes, code_width, crs, ccs, und, nullptr, stream));                   \
        break;
            RB_TC_DSUBS(X)
#undef X
        default: break;
        }
        if (getenv("RB_TC_STATS")) {  // debugging aid: how many (row, subquantizer) pairs the tensor pass left undecided
            std::vector<uint32_t> h(4 + regions);
            RB_CUDA_TRY(cudaMemcpyAsync(h.data(), work, h.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
            RB_CUDA_TRY(cudaStreamSynchronize(stream));
            size_t tot = 0;
            for (size_t r = 0; r < regions; r++) tot += h[4 + r];
            fprintf(stderr, "[rb tc] n=%zu M=%zu dsub=%zu: %zu of %zu pairs decided exactly among their candidates (%.3f%%)%s\n", n,
                    cb.M, cb.dsub, tot, n * cb.M, 100.0 * tot / (double)(n * cb.M), h[0] ? ", list overflow: exact kernel" : "");
        }
        RB_TRY(launch_encode_candidates(cb, x, ldx, und.cands, und.region_counts, (uint32_t)regions, (uint32_t)region_cap, codes,
                                        code_width, crs, ccs, stream));
        RB_TRY(launch_encode_exact_gated(cb, x, n, ldx, codes, code_width, crs, ccs, 0, und.overflow, 0u, stream));
        return RB_OK;
    };
    const rb_status st = body();
    cudaFreeAsync(work, stream);
    return st;
}

}  // namespace

rb_status launch_encode_tensor(const DeviceCodebook &cb, const TensorOperands &tc, const float *x, size_t n, ptrdiff_t ldx,
                               void *codes, int code_width, ptrdiff_t crs, ptrdiff_t ccs, cudaStream_t stream,
                               const RotatedInput *rot)
{
    if (n == 0) return RB_OK;
    if (!tc.ready() || !tensor_call_supported(cb, x, n, ldx)) {
        set_error("tensor encode path does not cover this call (k=%zu, dsub=%zu, ldx=%td)", cb.k, cb.dsub, ldx);
        return RB_ERR_UNSUPPORTED;
    }
    if (rot) {
        // A rotated input cannot fall back to the exact kernel on the same (approximate) buffer.  The flagged rows
        // are collected per subquantizer, [M] counters + [M][n] row indices, and re-decided by a kernel that first
        // re-rotates the subvector exactly (encode_exact.cu).
        const size_t units = n * cb.M;
        if (n > 0xffffffffull || !rotated_recheck_supported(cb, rot->d)) {
            set_error("tensor encode of a rotated batch does not cover this shape (n=%zu, d=%zu)", n, rot->d);
            return RB_ERR_UNSUPPORTED;
        }
        // undecided pairs first go to the candidate list (per-warp regions, a quarter of the pairs a warp can meet):
        // rotated_candidates_kernel decides them on the APPROXIMATE rotation whenever the candidates are further
        // apart than the rotation error can move them, and only the rest is bucketed per subquantizer ([M] counters +
        // [M][n] (row, candidate flags) pairs) for the exact re-rotation
        const size_t n_tiles = ceil_div(n, (size_t)kTile);
        const Plan plan = make_plan(cb.M, cb.dsub, n_tiles, device_sm_count(), 3);
        const size_t regions = (size_t)plan.ctas * 4 * kSets;
        const size_t region_cap = pairs_per_epilogue_warp(plan, cb.M, n_tiles) / 4 + 64;
        const size_t counters = (cb.M + 3) / 4 * 4, head = (regions + 3) / 4 * 4;
        uint32_t *work = nullptr;
        RB_CUDA_TRY(pool_malloc((void **)&work, (counters + head + 2 * units + 4 * regions * region_cap) * sizeof(uint32_t), stream));
        uint32_t *counts = work, *region_counts = work + counters, *cands = region_counts + head,
                 *rows = cands + 4 * regions * region_cap;
        Undecided und;
        und.bucket_counts = counts;
        und.bucket_rows = rows;
        und.cands = cands;
        und.region_counts = region_counts;
        und.region_cap = (uint32_t)region_cap;
        auto body = [&]() -> rb_status {
            RB_CUDA_TRY(cudaMemsetAsync(work, 0, (counters + head) * sizeof(uint32_t), stream));
            switch (cb.dsub) {
#define X(D)                                                                                                         \
    case D:                                                                                                          \
        RB_TRY(launch_t<D>(cb, tc, x, n, ldx, codes, code_width, crs, ccs, und, rot, stream));                       \
        break;
                RB_TC_DSUBS(X)
#undef X
            default: break;
            }
            RB_TRY(launch_rotated_candidates(cb, x, ldx, cands, region_counts, (uint32_t)regions, (uint32_t)region_cap, rot->rowerr,
                                             rot->sx_dev, rot->err_floor, counts, rows, n, codes, code_width, crs, ccs, stream));
            if (getenv("RB_TC_STATS")) {
                std::vector<uint32_t> h(counters + head);
                RB_CUDA_TRY(cudaMemcpyAsync(h.data(), work, h.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
                RB_CUDA_TRY(cudaStreamSynchronize(stream));
                size_t tot = 0, listed = 0;
                for (size_t m = 0; m < cb.M; m++) tot += h[m];
                for (size_t r = 0; r < regions; r++) listed += h[counters + r];
                fprintf(stderr, "[rb tc] rotated n=%zu M=%zu dsub=%zu: of %zu pairs %zu undecided by the tensor filter (%.3f%%), %zu "
                                "re-rotated exactly (%.4f%%)\n", n, cb.M, cb.dsub, units, listed, 100.0 * listed / (double)units, tot,
                        100.0 * tot / (double)units);
            }
            return launch_rotated_recheck(cb, counts, rows, n, rot->x0, rot->ldx0, rot->r, rot->d, codes, code_width, crs, ccs,
                                          stream);
        };
        const rb_status st = body();
        cudaFreeAsync(work, stream);
        return st;
    }
    // row chunks keep the pair counts (32 bits) far from wrapping: n_chunk * M < 2^31
    const size_t max_rows = (((size_t)1 << 31) - 1) / cb.M / kTile * kTile;
    for (size_t r0 = 0; r0 < n; r0 += max_rows) {
        const size_t rows = n - r0 < max_rows ? n - r0 : max_rows;
        RB_TRY(encode_tensor_chunk(cb, tc, x + (ptrdiff_t)r0 * ldx, rows, ldx,
                                   reinterpret_cast<char *>(codes) + (ptrdiff_t)r0 * crs * code_width, code_width, crs, ccs, stream));
    }
    return RB_OK;
}

}  // namespace rb
