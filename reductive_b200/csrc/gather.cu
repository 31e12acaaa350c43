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

// ---------------------------------------------------------------------------------------------------------
// Fast path (u8 codes, dense [n, M] code matrix, 16-byte aligned output rows, even dsub): tiled gather.
//
// A block owns one column group (a whole number of subquantizers whose width in floats is a multiple of 4; its
// codebook slice lives in shared memory for the block's lifetime) and walks a strip of row tiles.  Per tile the
// code bytes of the tile's rows are staged in shared memory by coalesced 16-byte loads that were issued one tile
// AHEAD (register prefetch), so the inner loop has no global-load latency in it: per 16-byte output piece it
// reads one or two code bytes and one 16-byte / two 8-byte centroid pieces from shared memory and issues one
// 16-byte global store.  Consecutive lanes own consecutive pieces of a row, so a warp store writes 512 contiguous
// bytes (row segments of the group).  Blocks of the column groups of one strip have adjacent block indices: they
// run concurrently and the 32-byte sectors straddling a group boundary are completed in L2.
// ---------------------------------------------------------------------------------------------------------
constexpr int kTileThreads = 512;
constexpr int kTileVpt = 2;  // 16-byte code vectors prefetched per thread and tile

// piece table entry: where the two 8-byte halves of a 16-byte output piece come from
struct PieceSrc {
    uint32_t off0, cc0, off1, cc1;  // float offset into the group's codebook (add code * dsub); code column
};

template <bool QUAD, bool CHECK>
__global__ void __launch_bounds__(kTileThreads, 2)
gather_tile_kernel(const float *__restrict__ quantizers, int k, int dsub, int M, const uint8_t *__restrict__ codes,
                   long long n, float *__restrict__ out, long long ldo, int m_per_group, int n_groups,
                   int tile_rows, long long tiles_per_strip, int cb_floats_max, int *__restrict__ err_flag)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int g = blockIdx.x % n_groups;
    const long long strip = blockIdx.x / n_groups;
    const int m0 = g * m_per_group;
    const int mg = min(m_per_group, M - m0);
    const int w4 = mg * dsub / 4;  // 16-byte pieces per row in this group

    float *cb = reinterpret_cast<float *>(smem_raw);                       // [mg][k][dsub]
    PieceSrc *tab = reinterpret_cast<PieceSrc *>(cb + cb_floats_max);      // [w4]
    uint8_t *sc = reinterpret_cast<uint8_t *>(tab + (m_per_group * dsub / 4));  // [tile_rows][M]

    {
        const float4 *src = reinterpret_cast<const float4 *>(quantizers + (size_t)m0 * k * dsub);
        float4 *dst = reinterpret_cast<float4 *>(cb);
        const int total4 = mg * k * dsub / 4;
        for (int i = threadIdx.x; i < total4; i += kTileThreads) dst[i] = __ldg(src + i);
        for (int c4 = threadIdx.x; c4 < w4; c4 += kTileThreads) {
            PieceSrc e;
            const int col0 = 4 * c4, col1 = 4 * c4 + 2;
            e.cc0 = col0 / dsub;
            e.off0 = e.cc0 * k * dsub + col0 % dsub;
            e.cc1 = col1 / dsub;
            e.off1 = e.cc1 * k * dsub + col1 % dsub;
            tab[c4] = e;
        }
    }

    const long long tile0 = strip * tiles_per_strip;
    const long long n_tiles = (n + tile_rows - 1) / tile_rows;
    const long long tile1 = min(n_tiles, tile0 + tiles_per_strip);
    const size_t total_code_bytes = (size_t)n * M;
    const int tile_vecs = tile_rows * M / 16;  // tile_rows * M is a multiple of 16 by construction

    // register prefetch of a tile's code bytes (vectors past the end of the code matrix are skipped; a ragged
    // last vector is read bytewise)
    uint4 pre[kTileVpt];
    auto prefetch = [&](long long tile) {
        const size_t base = (size_t)tile * tile_rows * M;
#pragma unroll
        for (int v = 0; v < kTileVpt; v++) {
            const int vi = threadIdx.x + v * kTileThreads;
            const size_t byte0 = base + (size_t)vi * 16;
            pre[v] = make_uint4(0u, 0u, 0u, 0u);
            if (vi < tile_vecs && byte0 < total_code_bytes) {
                if (byte0 + 16 <= total_code_bytes) {
                    pre[v] = __ldcs(reinterpret_cast<const uint4 *>(codes + byte0));
                } else {
                    uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                    for (int b = 0; b < 16; b++)
                        if (byte0 + b < total_code_bytes) w[b >> 2] |= (uint32_t)codes[byte0 + b] << (8 * (b & 3));
                    pre[v] = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
    };

    const int d_row = kTileThreads / w4, d_c = kTileThreads % w4;
    const int gcol0 = m0 * dsub;
    bool bad = false;

    if (tile0 < tile1) prefetch(tile0);
    for (long long tile = tile0; tile < tile1; tile++) {
        __syncthreads();  // everyone is done with the previous tile's codes (and, first time, the codebook is in)
#pragma unroll
        for (int v = 0; v < kTileVpt; v++) {
            const int vi = threadIdx.x + v * kTileThreads;
            if (vi < tile_vecs) reinterpret_cast<uint4 *>(sc)[vi] = pre[v];
        }
        __syncthreads();
        if (tile + 1 < tile1) prefetch(tile + 1);

        const long long r0 = tile * tile_rows;
        const int rows = (int)min((long long)tile_rows, n - r0);
        int row = (int)threadIdx.x / w4, c4 = (int)threadIdx.x % w4;
        float *orow = out + r0 * ldo + gcol0;
        while (row < rows) {
            const PieceSrc e = tab[c4];
            const uint8_t *crow = sc + row * M + m0;
            float4 v;
            if constexpr (QUAD) {  // dsub % 4 == 0: the piece lies inside one centroid
                unsigned c = crow[e.cc0];
                if constexpr (CHECK) {
                    bad |= c >= (unsigned)k;
                    c = min(c, (unsigned)(k - 1));
                }
                v = *reinterpret_cast<const float4 *>(cb + e.off0 + c * dsub);
            } else {
                unsigned ca = crow[e.cc0], cbb = crow[e.cc1];
                if constexpr (CHECK) {
                    bad |= (ca >= (unsigned)k) | (cbb >= (unsigned)k);
                    ca = min(ca, (unsigned)(k - 1));
                    cbb = min(cbb, (unsigned)(k - 1));
                }
                const float2 lo = *reinterpret_cast<const float2 *>(cb + e.off0 + ca * dsub);
                const float2 hi = *reinterpret_cast<const float2 *>(cb + e.off1 + cbb * dsub);
                v = make_float4(lo.x, lo.y, hi.x, hi.y);
            }
            __stcs(reinterpret_cast<float4 *>(orow + (long long)row * ldo + 4 * c4), v);
            row += d_row;
            c4 += d_c;
            if (c4 >= w4) {
                c4 -= w4;
                row++;
            }
        }
    }
    if (bad) atomicExch(err_flag, 1);
}

// Returns true when the tiled kernel was launched (otherwise the caller uses the generic kernel).
bool launch_tiled(const DeviceCodebook &cb, const void *codes, int code_width, size_t n, ptrdiff_t crs, ptrdiff_t ccs,
                  float *out, ptrdiff_t ldo, int *err_flag, cudaStream_t stream, rb_status *status)
{
    const int M = (int)cb.M, k = (int)cb.k, dsub = (int)cb.dsub;
    *status = RB_OK;
    if (code_width != 1 || ccs != 1 || crs != (ptrdiff_t)M || (dsub & 1) || M > 4096) return false;
    if ((reinterpret_cast<uintptr_t>(out) & 15) || (ldo & 3) || (reinterpret_cast<uintptr_t>(codes) & 15)) return false;
    if ((reinterpret_cast<uintptr_t>(cb.quantizers) & 15) || n < 256) return false;
    const size_t per_m = (size_t)k * dsub * sizeof(float);
    const size_t smem_budget = 113 * 1024;  // two 512-thread blocks per SM
    const size_t cb_budget = 100 * 1024;
    const int step = (dsub % 4 == 0) ? 1 : 2;  // group width must be a multiple of 4 floats
    int mg = (int)(cb_budget / per_m);
    mg -= mg % step;
    if (mg < step) return false;
    if (mg > M) mg = M;
    int n_groups = (int)ceil_div(M, mg);
    mg = (int)ceil_div(M, n_groups);
    mg += (step - mg % step) % step;
    n_groups = (int)ceil_div(M, mg);
    if (((M - (n_groups - 1) * mg) * dsub) % 4 != 0) return false;  // last group's width
    if ((size_t)mg * k * dsub % 4 != 0) return false;
    const size_t cb_bytes = (size_t)mg * per_m;
    const size_t tab_bytes = (size_t)(mg * dsub / 4) * sizeof(PieceSrc);
    if (cb_bytes + tab_bytes + 16 * (size_t)M > smem_budget) return false;
    size_t code_bytes = smem_budget - cb_bytes - tab_bytes;
    if (code_bytes > (size_t)kTileThreads * kTileVpt * 16) code_bytes = (size_t)kTileThreads * kTileVpt * 16;
    int tile_rows = (int)(code_bytes / M);
    tile_rows -= tile_rows % 16;
    if (tile_rows > 512) tile_rows = 512;
    if (tile_rows < 16) return false;
    const size_t smem = cb_bytes + tab_bytes + (size_t)tile_rows * M;

    const long long n_tiles = (long long)ceil_div(n, (size_t)tile_rows);
    long long strips = (2 * 148) / n_groups;  // one resident wave
    if (strips < 1) strips = 1;
    if (strips > n_tiles) strips = n_tiles;
    const long long tiles_per_strip = (n_tiles + strips - 1) / strips;
    strips = (n_tiles + tiles_per_strip - 1) / tiles_per_strip;
    const unsigned grid = (unsigned)(strips * n_groups);

    const bool quad = dsub % 4 == 0, check = k < 256;
    auto kern = quad ? (check ? gather_tile_kernel<true, true> : gather_tile_kernel<true, false>)
                     : (check ? gather_tile_kernel<false, true> : gather_tile_kernel<false, false>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        *status = RB_ERR_CUDA;
        return true;
    }
    kern<<<grid, kTileThreads, smem, stream>>>(cb.quantizers, k, dsub, M, reinterpret_cast<const uint8_t *>(codes),
                                              (long long)n, out, (long long)ldo, mg, n_groups, tile_rows,
                                              tiles_per_strip, (int)(cb_bytes / sizeof(float)), err_flag);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e), __FILE__, __LINE__);
        *status = RB_ERR_CUDA;
    }
    return true;
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
    rb_status st = RB_OK;
    if (launch_tiled(cb, codes, code_width, n, crs, ccs, out, ldo, err_flag, stream, &st)) return st;
    if (cb.dsub % 4 == 0) return launch_pw<4>(cb, codes, code_width, n, crs, ccs, out, ldo, err_flag, stream);
    if (cb.dsub % 2 == 0) return launch_pw<2>(cb, codes, code_width, n, crs, ccs, out, ldo, err_flag, stream);
    return launch_pw<1>(cb, codes, code_width, n, crs, ccs, out, ldo, err_flag, stream);
}

}  // namespace rb
