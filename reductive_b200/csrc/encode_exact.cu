// encode_exact.cu — exact FP32 SIMT nearest-centroid encode (sm_100a).
//
// Replaces, bit for bit, the reference chain
//   primitives::quantize_batch_into   src/pq/primitives.rs:64-104
//   -> kmeans::cluster_assignments    src/kmeans.rs:133-159
//   -> SquaredEuclideanDistance<Ix2>  src/linalg.rs:150-180
// by evaluating, per (row, subquantizer, centroid), the same floating-point expression tree:
//   xs  = unrolled_dot(x_sub, x_sub)                       (linalg.rs:167)
//   cs  = unrolled_dot(c_j, c_j)                           (linalg.rs:168, precomputed per codebook)
//   dp  = fma(x[K-1], c[K-1], ... fma(x[0], c[0], 0))      (linalg.rs:170, matrixmultiply FMA kernel,
//                                                           kc = 256 blocks combined by a plain add)
//   d_j = (xs + cs) - (dp + dp)                            (linalg.rs:173-174)
//   code = first j minimising d_j, NaN ranking largest     (kmeans.rs:150-155)
// without ever materialising the reference's n x k distance temporary.
//
// It is the correctness anchor, the path for shapes the tensor kernel does not cover, and (through
// launch_encode_recheck) the arbiter of near-ties the tensor kernel flags.
//
// Bound: FP32 pipe.  (dsub FMA + ~6 other ops) per (row, m, j) -> at C2 (2M x 300, M=30, k=256)
// ~2.5e11 lane-ops, i.e. >= 7 ms on 148 SMs x 128 lanes; HBM traffic is the algorithmic 4d+M bytes/row.
#include "common.cuh"

namespace rb {

namespace {

constexpr int kThreads = 128;
constexpr int kInvalid = -1;

// Load DSUB floats of one centroid from shared memory with the widest aligned vector loads.
template <int DSUB>
__device__ __forceinline__ void load_centroid(const float *__restrict__ src, float (&c)[DSUB])
{
    if constexpr (DSUB % 4 == 0) {
#pragma unroll
        for (int t = 0; t < DSUB; t += 4) {
            float4 v = *reinterpret_cast<const float4 *>(src + t);
            c[t] = v.x; c[t + 1] = v.y; c[t + 2] = v.z; c[t + 3] = v.w;
        }
    } else if constexpr (DSUB % 2 == 0) {
#pragma unroll
        for (int t = 0; t < DSUB; t += 2) {
            float2 v = *reinterpret_cast<const float2 *>(src + t);
            c[t] = v.x; c[t + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int t = 0; t < DSUB; t++) c[t] = src[t];
    }
}

// Full OrderedFloat semantics for one row (rare path: the fast loop found nothing below +inf).
template <int DSUB>
__device__ __noinline__ int slow_argmin(const float *__restrict__ quantizers_m, const float *__restrict__ cs_m,
                                        int k, const float (&x)[DSUB], float xs)
{
    int best = 0;
    float bv = 0.f;
    for (int j = 0; j < k; j++) {
        float dp = 0.f;
#pragma unroll
        for (int t = 0; t < DSUB; t++) dp = __fmaf_rn(x[t], quantizers_m[(size_t)j * DSUB + t], dp);
        float d = ref_distance(xs, cs_m[j], dp);
        if (j == 0 || of_less(d, bv)) {
            bv = d;
            best = j;
        }
    }
    return best;
}

// One block: kThreads*RPT rows x a range of subquantizers.  Centroids of the current subquantizer are
// staged in shared memory (in chunks of kch centroids) and broadcast to all lanes.
template <int DSUB, int RPT>
__global__ void __launch_bounds__(kThreads)
encode_exact_kernel(const float *__restrict__ quantizers, const float *__restrict__ cs_all, int M, int k,
                    int kch, const float *__restrict__ x, long long n, long long ldx, void *codes,
                    int code_width, long long crs, long long ccs, int seq_norm, int m_per_block,
                    const uint32_t *__restrict__ gate, uint32_t gate_thr)
{
    extern __shared__ __align__(16) float smem[];
    if (gate != nullptr && *gate <= gate_thr) return;  // gated fallback of the tensor path: nothing overflowed
    float *cen = smem;                      // [kch][DSUB]
    float *csm = smem + (size_t)kch * DSUB; // [kch]

    const long long tile_base = (long long)blockIdx.x * (kThreads * RPT);
    const int m0 = blockIdx.y * m_per_block;
    const int m1 = min(M, m0 + m_per_block);

    for (int m = m0; m < m1; m++) {
        const float *qm = quantizers + (size_t)m * k * DSUB;
        const float *csg = cs_all + (size_t)m * k;

        float xv[RPT][DSUB];
        float xs[RPT];
        float best[RPT];
        int bidx[RPT];
#pragma unroll
        for (int r = 0; r < RPT; r++) {
            long long row = tile_base + (long long)r * kThreads + threadIdx.x;
            if (row < n) {
                const float *xr = x + row * ldx + (long long)m * DSUB;
#pragma unroll
                for (int t = 0; t < DSUB; t++) xv[r][t] = __ldg(xr + t);
            } else {
#pragma unroll
                for (int t = 0; t < DSUB; t++) xv[r][t] = 0.f;
            }
            xs[r] = seq_norm ? sequential_sqnorm_reg<DSUB>(xv[r]) : unrolled_sqnorm_reg<DSUB>(xv[r]);
            best[r] = __int_as_float(0x7f800000);  // +inf
            bidx[r] = kInvalid;
        }

        for (int j0 = 0; j0 < k; j0 += kch) {
            const int jn = min(kch, k - j0);
            __syncthreads();  // previous chunk fully consumed
            for (int i = threadIdx.x; i < jn * DSUB; i += kThreads) cen[i] = __ldg(qm + (size_t)j0 * DSUB + i);
            for (int i = threadIdx.x; i < jn; i += kThreads) csm[i] = __ldg(csg + j0 + i);
            __syncthreads();

#pragma unroll 2
            for (int j = 0; j < jn; j++) {
                float c[DSUB];
                load_centroid<DSUB>(cen + (size_t)j * DSUB, c);
                const float csj = csm[j];
#pragma unroll
                for (int r = 0; r < RPT; r++) {
                    float dp = 0.f;
#pragma unroll
                    for (int t = 0; t < DSUB; t++) dp = __fmaf_rn(xv[r][t], c[t], dp);
                    const float d = ref_distance(xs[r], csj, dp);
                    // strict '<' keeps the first minimum; NaN never wins here (handled below)
                    if (d < best[r]) {
                        best[r] = d;
                        bidx[r] = j0 + j;
                    }
                }
            }
        }

#pragma unroll
        for (int r = 0; r < RPT; r++) {
            long long row = tile_base + (long long)r * kThreads + threadIdx.x;
            if (row < n) {
                int idx = bidx[r];
                if (idx == kInvalid) idx = slow_argmin<DSUB>(qm, csg, k, xv[r], xs[r]);
                store_code(codes, code_width, row * crs + (long long)m * ccs, (unsigned)idx);
            }
        }
    }
}

// Any dsub (including > 256, where the kc = 256 blocking of matrixmultiply matters).  One thread per row,
// x re-read through L1; slow, only for shapes outside the templated set.
__global__ void __launch_bounds__(kThreads)
encode_exact_generic_kernel(const float *__restrict__ quantizers, const float *__restrict__ cs_all, int M, int k,
                            int dsub, const float *__restrict__ x, long long n, long long ldx, void *codes,
                            int code_width, long long crs, long long ccs, int seq_norm, int m_per_block,
                            const uint32_t *__restrict__ gate, uint32_t gate_thr)
{
    if (gate != nullptr && *gate <= gate_thr) return;
    const long long row = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (row >= n) return;
    const int m0 = blockIdx.y * m_per_block;
    const int m1 = min(M, m0 + m_per_block);
    for (int m = m0; m < m1; m++) {
        const float *qm = quantizers + (size_t)m * k * dsub;
        const float *csg = cs_all + (size_t)m * k;
        const float *xr = x + row * ldx + (long long)m * dsub;
        float xs;
        if (seq_norm) {
            xs = 0.f;
            for (int t = 0; t < dsub; t++) xs = __fadd_rn(xs, __fmul_rn(__ldg(xr + t), __ldg(xr + t)));
        } else {
            xs = unrolled_dot_dev(dsub, [&](int i) { return __ldg(xr + i); }, [&](int i) { return __ldg(xr + i); });
        }
        int best = 0;
        float bv = 0.f;
        for (int j = 0; j < k; j++) {
            const float *cj = qm + (size_t)j * dsub;
            float total = 0.f, acc = 0.f;
            bool have = false;
            for (int t = 0; t < dsub; t++) {
                acc = __fmaf_rn(__ldg(xr + t), __ldg(cj + t), acc);
                if ((t & 255) == 255 || t == dsub - 1) {  // matrixmultiply kc = 256: C = AB, then C = C + AB
                    total = have ? __fadd_rn(total, acc) : acc;
                    have = true;
                    acc = 0.f;
                }
            }
            const float d = ref_distance(xs, __ldg(csg + j), total);
            if (j == 0 || of_less(d, bv)) {
                bv = d;
                best = j;
            }
        }
        store_code(codes, code_width, row * crs + (long long)m * ccs, (unsigned)best);
    }
}

__global__ void centroid_norms_kernel(const float *__restrict__ q, size_t rows, int dsub, float *__restrict__ cs)
{
    size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float *c = q + r * dsub;
    cs[r] = unrolled_dot_dev(dsub, [&](int i) { return c[i]; }, [&](int i) { return c[i]; });
}

// One warp per flagged (row, m) pair: lanes stride over centroids with the full reference semantics.
__global__ void __launch_bounds__(256)
encode_recheck_kernel(const float *__restrict__ quantizers, const float *__restrict__ cs_all, int k, int dsub,
                      const float *__restrict__ x, long long ldx, const uint32_t *__restrict__ pairs,
                      const uint32_t *__restrict__ n_pairs_ptr, uint32_t max_pairs, void *codes, int code_width,
                      long long crs, long long ccs)
{
    const uint32_t n_pairs = min(*n_pairs_ptr, max_pairs);
    const int lane = threadIdx.x & 31;
    const uint32_t warps_per_grid = gridDim.x * (blockDim.x >> 5);
    for (uint32_t p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); p < n_pairs; p += warps_per_grid) {
        const long long row = pairs[2 * (size_t)p];
        const int m = (int)pairs[2 * (size_t)p + 1];
        const float *xr = x + row * ldx + (long long)m * dsub;
        const float *qm = quantizers + (size_t)m * k * dsub;
        const float *csg = cs_all + (size_t)m * k;
        const float xs = unrolled_dot_dev(dsub, [&](int i) { return __ldg(xr + i); }, [&](int i) { return __ldg(xr + i); });
        int best = -1;
        float bv = 0.f;
        for (int j = lane; j < k; j += 32) {
            const float *cj = qm + (size_t)j * dsub;
            float total = 0.f, acc = 0.f;
            bool have = false;
            for (int t = 0; t < dsub; t++) {
                acc = __fmaf_rn(__ldg(xr + t), __ldg(cj + t), acc);
                if ((t & 255) == 255 || t == dsub - 1) {
                    total = have ? __fadd_rn(total, acc) : acc;
                    have = true;
                    acc = 0.f;
                }
            }
            const float d = ref_distance(xs, __ldg(csg + j), total);
            if (best < 0 || of_less(d, bv)) {
                bv = d;
                best = j;
            }
        }
        // warp argmin: smaller distance wins, equal distance -> smaller index (first minimum)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
            const int oi = __shfl_xor_sync(0xffffffffu, best, off);
            const bool take = (oi >= 0) && (best < 0 || of_less(ov, bv) || (!of_less(bv, ov) && oi < best));
            if (take) {
                bv = ov;
                best = oi;
            }
        }
        if (lane == 0) store_code(codes, code_width, row * crs + (long long)m * ccs, (unsigned)best);
    }
}

// One thread per undecided (row, m) pair of the tensor filter: the candidates are the centroids j = 16 b + a for the
// flagged chains a and blocks b (bit layout: common.cuh launch_encode_candidates; every other centroid is proven not
// to be the reference's argmin; all ones = no proof, every centroid).  Same expression tree as encode_exact_kernel; smaller distance wins,
// equal distances -> smaller index, NaN ranks largest (kmeans.rs:150-155).  The list comes in regions (one per
// epilogue warp of the tensor kernel): a block walks whole regions.
template <int DSUB>
__global__ void __launch_bounds__(256)
encode_candidates_kernel(const float *__restrict__ quantizers, const float *__restrict__ cs_all, int k,
                         const float *__restrict__ x, long long ldx, const uint32_t *__restrict__ cands,
                         const uint32_t *__restrict__ region_counts, uint32_t regions, uint32_t region_cap, void *codes,
                         int code_width, long long crs, long long ccs)
{
    for (uint32_t r = blockIdx.x; r < regions; r += gridDim.x) {
      const uint32_t cnt = min(region_counts[r], region_cap);
      for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) {
        const uint4 e = __ldg(reinterpret_cast<const uint4 *>(cands) + (size_t)r * region_cap + i);
        const long long row = e.x;
        const int m = (int)e.y;
        const uint32_t ma = e.z, mb = (e.w | (e.w >> 16)) & 0xffffu;
        const float *xr = x + row * ldx + (long long)m * DSUB;
        float xv[DSUB];
        if constexpr (DSUB % 2 == 0) {  // the tensor path requires ldx % 4 == 0 and even dsub: 8-byte aligned
#pragma unroll
            for (int t = 0; t < DSUB; t += 2) {
                const float2 v = __ldg(reinterpret_cast<const float2 *>(xr + t));
                xv[t] = v.x;
                xv[t + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int t = 0; t < DSUB; t++) xv[t] = __ldg(xr + t);
        }
        const float xs = unrolled_sqnorm_reg<DSUB>(xv);
        const float *qm = quantizers + (size_t)m * k * DSUB;
        const float *csg = cs_all + (size_t)m * k;
        int best = -1;
        float bv = 0.f;
        for (uint32_t bb = mb; bb; bb &= bb - 1) {  // ascending blocks, ascending chains: ascending j
            const int b = __ffs(bb) - 1;
            for (int a = 0; a < 16; a++) {
                if (!((ma >> ((a >> 1) + ((a & 1) << 4))) & 1u)) continue;
                const int j = 16 * b + a;
                if (j >= k) continue;
                float c[DSUB];
                load_centroid<DSUB>(qm + (size_t)j * DSUB, c);
                float dp = 0.f;
#pragma unroll
                for (int t = 0; t < DSUB; t++) dp = __fmaf_rn(xv[t], c[t], dp);
                const float d = ref_distance(xs, __ldg(csg + j), dp);
                if (best < 0 || of_less(d, bv)) {  // j ascends: strict "less" keeps the first minimum
                    bv = d;
                    best = j;
                }
            }
        }
        if (best >= 0) store_code(codes, code_width, row * crs + (long long)m * ccs, (unsigned)best);
      }
    }
}

// Candidate lists of a ROTATED batch.  x holds the approximate rotation A of the rows (|A - T| <= ds per subvector in
// the 2-norm, T the reference's exact rotation); the reference decides argmin_j tree(T, c_j).  With D_j = tree(A, c_j):
//   | ||T - c_j||^2 - ||A - c_j||^2 | = |(T - A).(T + A - 2 c_j)| <= ds (2 ||A - c_j|| + ds)
// and tree(v, c) is within e_fl = 2^-20 (||v||^2 + ||c||^2) of the real ||v - c||^2 (FP32 roundings of the norms, the
// FMA chain and the two adds; generous), so the reference's value for j lies within
//   E_j = ds (2 sqrt(max(D_j, 0) + e_fl) + ds) + 2 e_fl   of D_j.
// If the best candidate b has D_b + E_b < D_j - E_j for every other candidate j, b is the reference's argmin (no tie
// possible); otherwise the row goes to the exact re-rotation.
template <int DSUB>
__global__ void __launch_bounds__(256)
rotated_candidates_kernel(const float *__restrict__ quantizers, const float *__restrict__ cs_all, int k,
                          const float *__restrict__ x, long long ldx, const uint32_t *__restrict__ cands,
                          const uint32_t *__restrict__ region_counts, uint32_t regions, uint32_t region_cap,
                          const float *__restrict__ rowerr, const float *__restrict__ sx_dev, float err_floor,
                          uint32_t *__restrict__ bucket_counts, uint32_t *__restrict__ bucket_rows, long long n_cap,
                          void *codes, int code_width, long long crs, long long ccs)
{
    const float floor_err = err_floor / sx_dev[0];
    for (uint32_t r = blockIdx.x; r < regions; r += gridDim.x) {
      const uint32_t cnt = min(region_counts[r], region_cap);
      for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) {
        const uint4 e = __ldg(reinterpret_cast<const uint4 *>(cands) + (size_t)r * region_cap + i);
        const long long row = e.x;
        const int m = (int)e.y;
        const uint32_t ma = e.z, mb = (e.w | (e.w >> 16)) & 0xffffu;
        const float perr = rowerr[row] + floor_err;
        bool decided = false;
        int best = -1;
        if (e.z != 0xffffffffu && perr < 3.0e38f) {  // (all ones: no proof from the filter; NaN: row not rotated approximately)
            const float *xr = x + row * ldx + (long long)m * DSUB;
            float xv[DSUB];
#pragma unroll
            for (int t = 0; t < DSUB; t += 2) {
                const float2 v = __ldg(reinterpret_cast<const float2 *>(xr + t));
                xv[t] = v.x;
                xv[t + 1] = v.y;
            }
            const float xs = unrolled_sqnorm_reg<DSUB>(xv);
            const float ds = sqrtf((float)DSUB) * perr * 1.02f;
            const float *qm = quantizers + (size_t)m * k * DSUB;
            const float *csg = cs_all + (size_t)m * k;
            float d1 = __int_as_float(0x7f800000), lo2 = __int_as_float(0x7f800000);  // best D, smallest D_j - E_j of the others
            float e1 = 0.f;
            for (uint32_t bb = mb; bb; bb &= bb - 1) {
                const int b = __ffs(bb) - 1;
                for (int a = 0; a < 16; a++) {
                    if (!((ma >> ((a >> 1) + ((a & 1) << 4))) & 1u)) continue;
                    const int j = 16 * b + a;
                    if (j >= k) continue;
                    float c[DSUB];
                    load_centroid<DSUB>(qm + (size_t)j * DSUB, c);
                    float dp = 0.f;
#pragma unroll
                    for (int t = 0; t < DSUB; t++) dp = __fmaf_rn(xv[t], c[t], dp);
                    const float csj = __ldg(csg + j);
                    const float d = ref_distance(xs, csj, dp);
                    const float efl = 9.5367431640625e-7f * (xs + csj);
                    const float ej = fmaf(ds, fmaf(2.f, sqrtf(fmaxf(d, 0.f) + efl), ds), 2.f * efl);
                    if (d < d1) {  // the previous best becomes one of the others
                        lo2 = fminf(lo2, d1 - e1);
                        d1 = d;
                        e1 = ej;
                        best = j;
                    } else {
                        lo2 = fminf(lo2, d - ej);
                    }
                }
            }
            decided = best >= 0 && (d1 + e1 < lo2) && (d1 == d1) && (e1 < 3.0e38f);
        }
        if (decided) store_code(codes, code_width, row * crs + (long long)m * ccs, (unsigned)best);
        // Undecided pairs go to their subquantizer's bucket.  One atomic per (warp, subquantizer): the lanes that
        // append to the same bucket elect a leader (a per-lane atomicAdd on M counters serialised the whole kernel).
        const unsigned active = __activemask();
        const unsigned und = __ballot_sync(active, !decided);
        if (!decided) {
            const unsigned peers = __match_any_sync(und, m);
            const int leader = __ffs(peers) - 1;
            uint32_t base = 0;
            if ((int)(threadIdx.x & 31) == leader) base = atomicAdd(&bucket_counts[m], (uint32_t)__popc(peers));
            base = __shfl_sync(peers, base, leader);
            const uint32_t slot = base + __popc(peers & ((1u << (threadIdx.x & 31)) - 1u));  // slot < n_cap: once per (row, m)
            // chain flags in natural order (bit a = chain a) | block flags << 16; all ones: every centroid
            uint32_t fl = 0xffffffffu;
            if (e.z != 0xffffffffu) {
                uint32_t chains = 0u;
                for (int a = 0; a < 16; a++) chains |= ((ma >> ((a >> 1) + ((a & 1) << 4))) & 1u) << a;
                fl = chains | (mb << 16);
            }
            reinterpret_cast<uint2 *>(bucket_rows)[(size_t)m * (size_t)n_cap + slot] = make_uint2((uint32_t)row, fl);
        }
      }
    }
}

// Flagged rows of a rotated batch (tensor rotation + tensor encode, encode_tc.cuh RotatedInput), bucketed per
// subquantizer.  blockIdx.y = subquantizer m: the block keeps R[:, m*DSUB .. +DSUB) and the m-th codebook in shared
// memory and walks its share of the bucket, thread = flagged row:
//   rx_sub = x0_row . R[:, cols]   one FMA chain per component over K blocks of 256 combined by a plain add
//                                  (matrixmultiply's order, as project.cu) -> the bits the reference's rx holds
//   code   = the reference's argmin over the k centroids (same expression tree as encode_exact_kernel)
template <int DSUB>
__global__ void __launch_bounds__(256)
rotated_recheck_kernel(const uint32_t *__restrict__ counts, const uint32_t *__restrict__ rows, long long n_cap,
                       const float *__restrict__ x0, long long ldx0, const float *__restrict__ r, int d,
                       const float *__restrict__ quantizers, const float *__restrict__ cs_all, int k, void *codes,
                       int code_width, long long crs, long long ccs)
{
    extern __shared__ __align__(16) float smem[];
    const int m = blockIdx.y;
    const uint32_t cnt = (uint32_t)min((long long)counts[m], n_cap);
    if ((uint32_t)blockIdx.x * 256u * (DSUB <= 16 ? 2u : 1u) >= cnt) return;
    float *rs = smem;                    // [d][DSUB]
    float *cen = rs + (size_t)d * DSUB;  // [k][DSUB]
    float *csm = cen + (size_t)k * DSUB; // [k]
    const float *qm = quantizers + (size_t)m * k * DSUB;
    const float *csg = cs_all + (size_t)m * k;
    for (int i = threadIdx.x; i < d * DSUB; i += 256) {
        const int ii = i / DSUB, t = i - ii * DSUB;
        rs[i] = __ldg(r + (size_t)ii * d + m * DSUB + t);
    }
    for (int i = threadIdx.x; i < k * DSUB; i += 256) cen[i] = __ldg(qm + i);
    for (int i = threadIdx.x; i < k; i += 256) csm[i] = __ldg(csg + i);
    __syncthreads();
    // RPT flagged rows per thread: every shared-memory operand read (a row of R, a centroid) feeds RPT FMA chains
    constexpr int RPT = DSUB <= 16 ? 2 : 1;
    for (uint32_t base = blockIdx.x * (256u * RPT); base < cnt; base += gridDim.x * (256u * RPT)) {
        long long row[RPT];
        uint32_t flags[RPT];
        const float *xr[RPT];
        bool live[RPT];
#pragma unroll
        for (int q = 0; q < RPT; q++) {
            const uint32_t idx = base + q * 256u + threadIdx.x;
            live[q] = idx < cnt;
            const uint2 ent = reinterpret_cast<const uint2 *>(rows)[(size_t)m * (size_t)n_cap + (live[q] ? idx : base)];  // base < cnt
            row[q] = ent.x;
            flags[q] = ent.y;
            xr[q] = x0 + row[q] * ldx0;
        }
        // One pass over the row in steps of four components (d % 4 == 0, 16-byte aligned rows:
        // project_tensor_call_supported), the next step's values already in flight; at every K block boundary of 256
        // the block's chain is folded into the running result by a plain add (first block: assignment).
        float rx[RPT][DSUB], acc[RPT][DSUB];
        float4 nxt[RPT];
#pragma unroll
        for (int q = 0; q < RPT; q++) {
            nxt[q] = __ldg(reinterpret_cast<const float4 *>(xr[q]));
#pragma unroll
            for (int t = 0; t < DSUB; t++) acc[q][t] = rx[q][t] = 0.f;
        }
        for (int i = 0; i < d; i += 4) {
            float xe[RPT][4];
#pragma unroll
            for (int q = 0; q < RPT; q++) {
                xe[q][0] = nxt[q].x; xe[q][1] = nxt[q].y; xe[q][2] = nxt[q].z; xe[q][3] = nxt[q].w;
                nxt[q] = __ldg(reinterpret_cast<const float4 *>(xr[q] + min(i + 4, d - 4)));
            }
#pragma unroll
            for (int e = 0; e < 4; e++) {
                float c[DSUB];
                load_centroid<DSUB>(rs + (size_t)(i + e) * DSUB, c);
#pragma unroll
                for (int q = 0; q < RPT; q++)
#pragma unroll
                    for (int t = 0; t < DSUB; t++) acc[q][t] = __fmaf_rn(xe[q][e], c[t], acc[q][t]);
            }
            if (((i + 4) & 255) == 0 || i + 4 >= d) {
                const bool first = i < 256;
#pragma unroll
                for (int q = 0; q < RPT; q++)
#pragma unroll
                    for (int t = 0; t < DSUB; t++) {
                        rx[q][t] = first ? acc[q][t] : __fadd_rn(rx[q][t], acc[q][t]);
                        acc[q][t] = 0.f;
                    }
            }
        }
        float xs[RPT], best[RPT];
        int bidx[RPT];
#pragma unroll
        for (int q = 0; q < RPT; q++) {
            xs[q] = unrolled_sqnorm_reg<DSUB>(rx[q]);
            best[q] = __int_as_float(0x7f800000);
            bidx[q] = kInvalid;
        }
        // only the centroids the tensor filter could not rule out (chain a = j mod 16 and block b = j / 16 both flagged;
        // all ones: every centroid), in ascending j: strict '<' keeps the first minimum
#pragma unroll
        for (int q = 0; q < RPT; q++) {
            const uint32_t chains = flags[q] & 0xffffu, blocks = flags[q] >> 16;
            for (uint32_t bb = blocks; bb; bb &= bb - 1) {
                const int b = __ffs(bb) - 1;
                for (uint32_t aa = chains; aa; aa &= aa - 1) {
                    const int j = 16 * b + __ffs(aa) - 1;
                    if (j >= k) continue;
                    float c[DSUB];
                    load_centroid<DSUB>(cen + (size_t)j * DSUB, c);
                    float dp = 0.f;
#pragma unroll
                    for (int t = 0; t < DSUB; t++) dp = __fmaf_rn(rx[q][t], c[t], dp);
                    const float dist = ref_distance(xs[q], csm[j], dp);
                    if (dist < best[q]) {
                        best[q] = dist;
                        bidx[q] = j;
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < RPT; q++) {
            if (!live[q]) continue;
            int bi = bidx[q];
            if (bi == kInvalid) bi = slow_argmin<DSUB>(qm, csg, k, rx[q], xs[q]);
            store_code(codes, code_width, row[q] * crs + (long long)m * ccs, (unsigned)bi);
        }
    }
}

template <int DSUB>
rb_status launch_rotated_t(const DeviceCodebook &cb, const uint32_t *counts, const uint32_t *rows, size_t n_cap,
                           const float *x0, ptrdiff_t ldx0, const float *r, size_t d, void *codes, int code_width,
                           ptrdiff_t crs, ptrdiff_t ccs, cudaStream_t stream)
{
    const size_t smem = (d * DSUB + cb.k * DSUB + cb.k) * sizeof(float);
    auto kern = rotated_recheck_kernel<DSUB>;
    if (smem > 48 * 1024) RB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // a few blocks per subquantizer, each staging its operands once; about four blocks' worth of work per SM
    unsigned per_m = (unsigned)ceil_div((size_t)sm_count() * 4, cb.M);
    const unsigned most = (unsigned)ceil_div(n_cap, (size_t)256);
    if (per_m > most) per_m = most;
    if (per_m < 1) per_m = 1;
    kern<<<dim3(per_m, (unsigned)cb.M), 256, smem, stream>>>(counts, rows, (long long)n_cap, x0, (long long)ldx0, r, (int)d,
                                                             cb.quantizers, cb.cs, (int)cb.k, codes, code_width,
                                                             (long long)crs, (long long)ccs);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

template <int DSUB>
rb_status launch_t(const DeviceCodebook &cb, const float *x, size_t n, ptrdiff_t ldx, void *codes, int code_width,
                   ptrdiff_t crs, ptrdiff_t ccs, int seq_norm, const uint32_t *gate, uint32_t gate_thr, cudaStream_t stream)
{
    constexpr int RPT = DSUB <= 16 ? 4 : 2;
    const int k = (int)cb.k, M = (int)cb.M;
    // centroid chunk so that shared memory stays <= 48 KB (no opt-in needed, several blocks per SM)
    int kch = (48 * 1024) / ((DSUB + 1) * (int)sizeof(float));
    kch = kch > k ? k : (kch / 8) * 8;
    const size_t smem = (size_t)kch * (DSUB + 1) * sizeof(float);
    const size_t row_tiles = ceil_div(n, (size_t)kThreads * RPT);
    // enough blocks to fill the SMs a few times over: split the subquantizers when there are few row tiles
    int m_per_block = M;
    const int sms = sm_count();
    while (m_per_block > 1 && row_tiles * ceil_div(M, m_per_block) < (size_t)sms * 4) m_per_block = (m_per_block + 1) / 2;
    dim3 grid((unsigned)row_tiles, (unsigned)ceil_div(M, m_per_block));
    encode_exact_kernel<DSUB, RPT><<<grid, kThreads, smem, stream>>>(
        cb.quantizers, cb.cs, M, k, kch, x, (long long)n, (long long)ldx, codes, code_width, (long long)crs,
        (long long)ccs, seq_norm, m_per_block, gate, gate_thr);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

}  // namespace

rb_status launch_encode_exact(const DeviceCodebook &cb, const float *x, size_t n, ptrdiff_t ldx, void *codes,
                              int code_width, ptrdiff_t crs, ptrdiff_t ccs, int seq_norm, cudaStream_t stream)
{
    return launch_encode_exact_gated(cb, x, n, ldx, codes, code_width, crs, ccs, seq_norm, nullptr, 0, stream);
}

rb_status launch_encode_exact_gated(const DeviceCodebook &cb, const float *x, size_t n, ptrdiff_t ldx, void *codes,
                                    int code_width, ptrdiff_t crs, ptrdiff_t ccs, int seq_norm, const uint32_t *gate,
                                    uint32_t gate_thr, cudaStream_t stream)
{
    if (n == 0) return RB_OK;
    if (cb.k > (size_t)INT32_MAX || cb.M > (size_t)INT32_MAX || cb.dsub > (size_t)INT32_MAX) {
        set_error("codebook extents exceed 2^31");
        return RB_ERR_UNSUPPORTED;
    }
#define RB_CASE(D) \
    case D: return launch_t<D>(cb, x, n, ldx, codes, code_width, crs, ccs, seq_norm, gate, gate_thr, stream);
    switch (cb.dsub) {
        RB_CASE(1) RB_CASE(2) RB_CASE(3) RB_CASE(4) RB_CASE(5) RB_CASE(6) RB_CASE(7) RB_CASE(8)
        RB_CASE(9) RB_CASE(10) RB_CASE(12) RB_CASE(15) RB_CASE(16) RB_CASE(20) RB_CASE(24) RB_CASE(25)
        RB_CASE(30) RB_CASE(32)
    default: break;
    }
#undef RB_CASE
    const int M = (int)cb.M;
    const size_t row_tiles = ceil_div(n, (size_t)kThreads);
    int m_per_block = M;
    const int sms = sm_count();
    while (m_per_block > 1 && row_tiles * ceil_div(M, m_per_block) < (size_t)sms * 4) m_per_block = (m_per_block + 1) / 2;
    dim3 grid((unsigned)row_tiles, (unsigned)ceil_div(M, m_per_block));
    encode_exact_generic_kernel<<<grid, kThreads, 0, stream>>>(cb.quantizers, cb.cs, M, (int)cb.k, (int)cb.dsub, x,
                                                             (long long)n, (long long)ldx, codes, code_width,
                                                             (long long)crs, (long long)ccs, seq_norm, m_per_block, gate,
                                                             gate_thr);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

rb_status launch_centroid_norms(const float *quantizers, size_t rows, size_t dsub, float *cs, cudaStream_t stream)
{
    if (rows == 0) return RB_OK;
    centroid_norms_kernel<<<(unsigned)ceil_div(rows, 128), 128, 0, stream>>>(quantizers, rows, (int)dsub, cs);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

// the subvector widths the tensor encode is instantiated for (encode_tc.cu RB_TC_DSUBS)
#define RB_ROT_DSUBS(X) X(2) X(4) X(6) X(8) X(10) X(12) X(16) X(20) X(24) X(30) X(32)

bool rotated_recheck_supported(const DeviceCodebook &cb, size_t d)
{
    if (cb.M > 65535 || d != cb.M * cb.dsub) return false;
    if ((d * cb.dsub + cb.k * cb.dsub + cb.k) * sizeof(float) > 200 * 1024) return false;
    switch (cb.dsub) {
#define X(D) case D:
        RB_ROT_DSUBS(X)
#undef X
        return true;
    default: return false;
    }
}

rb_status launch_rotated_recheck(const DeviceCodebook &cb, const uint32_t *counts, const uint32_t *rows, size_t n_cap,
                                 const float *x0, ptrdiff_t ldx0, const float *r, size_t d, void *codes, int code_width,
                                 ptrdiff_t crs, ptrdiff_t ccs, cudaStream_t stream)
{
    switch (cb.dsub) {
#define X(D)                                                                                                             \
    case D:                                                                                                              \
        return launch_rotated_t<D>(cb, counts, rows, n_cap, x0, ldx0, r, d, codes, code_width, crs, ccs, stream);
        RB_ROT_DSUBS(X)
#undef X
    default: break;
    }
    set_error("rotated recheck: subvector width %zu is not instantiated", cb.dsub);
    return RB_ERR_UNSUPPORTED;
}

rb_status launch_encode_candidates(const DeviceCodebook &cb, const float *x, ptrdiff_t ldx, const uint32_t *cands,
                                   const uint32_t *region_counts, uint32_t regions, uint32_t region_cap, void *codes,
                                   int code_width, ptrdiff_t crs, ptrdiff_t ccs, cudaStream_t stream)
{
    if (regions == 0) return RB_OK;
    const unsigned blocks = regions;  // one block per region (about 12 per SM): an even share for every block
    switch (cb.dsub) {
#define X(D)                                                                                                             \
    case D:                                                                                                              \
        encode_candidates_kernel<D><<<blocks, 128, 0, stream>>>(cb.quantizers, cb.cs, (int)cb.k, x, (long long)ldx, cands, \
                                                                region_counts, regions, region_cap, codes, code_width,  \
                                                                (long long)crs, (long long)ccs);                        \
        break;
        RB_ROT_DSUBS(X)
#undef X
    default:
        set_error("candidate recheck: subvector width %zu is not instantiated", cb.dsub);
        return RB_ERR_UNSUPPORTED;
    }
    RB_LAUNCH_CHECK();
    return RB_OK;
}

rb_status launch_rotated_candidates(const DeviceCodebook &cb, const float *x, ptrdiff_t ldx, const uint32_t *cands,
                                    const uint32_t *region_counts, uint32_t regions, uint32_t region_cap,
                                    const float *rowerr, const float *sx_dev, float err_floor, uint32_t *bucket_counts,
                                    uint32_t *bucket_rows, size_t n_cap, void *codes, int code_width, ptrdiff_t crs,
                                    ptrdiff_t ccs, cudaStream_t stream)
{
    if (regions == 0) return RB_OK;
    const unsigned blocks = regions;
    switch (cb.dsub) {
#define X(D)                                                                                                             \
    case D:                                                                                                              \
        rotated_candidates_kernel<D><<<blocks, 128, 0, stream>>>(cb.quantizers, cb.cs, (int)cb.k, x, (long long)ldx, cands, \
                                                                 region_counts, regions, region_cap, rowerr, sx_dev,     \
                                                                 err_floor, bucket_counts, bucket_rows, (long long)n_cap, \
                                                                 codes, code_width, (long long)crs, (long long)ccs);    \
        break;
        RB_ROT_DSUBS(X)
#undef X
    default:
        set_error("rotated candidate recheck: subvector width %zu is not instantiated", cb.dsub);
        return RB_ERR_UNSUPPORTED;
    }
    RB_LAUNCH_CHECK();
    return RB_OK;
}

rb_status launch_encode_recheck(const DeviceCodebook &cb, const float *x, ptrdiff_t ldx, const uint32_t *pairs,
                                const uint32_t *n_pairs, uint32_t max_pairs, void *codes, int code_width,
                                ptrdiff_t crs, ptrdiff_t ccs, cudaStream_t stream)
{
    if (max_pairs == 0) return RB_OK;
    encode_recheck_kernel<<<(unsigned)sm_count() * 4, 256, 0, stream>>>(cb.quantizers, cb.cs, (int)cb.k, (int)cb.dsub, x,
                                                        (long long)ldx, pairs, n_pairs, max_pairs, codes, code_width,
                                                        (long long)crs, (long long)ccs);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

}  // namespace rb
