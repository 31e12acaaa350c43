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
#include "sm100_ptx.cuh"

#include <cstdlib>

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
// Fast path (u8 codes, dense [n, M] code matrix, 16-byte aligned output rows, even dsub, d % 4 == 0): tiled gather.
//
// A block owns one column group — `gw` consecutive output columns, not necessarily whole subquantizers — and walks
// a strip of row tiles.  Its slice of the codebook lives in shared memory for the block's lifetime in a
// COLUMN-ALIGNED layout: cbt[c][j] = value of output column col_lo + j when that column's subquantizer has code c,
// row pitch W a multiple of 32 floats.  The bank of cbt[c][j] is then j mod 32 whatever c is, so the eight lanes
// of a quarter warp, which own consecutive 16-byte pieces of one row, read 32 distinct banks: the scattered
// centroid reads are conflict free by construction (the natural [m][c][dsub] layout replayed 58 % of them).
// The code bytes of a tile's rows are contiguous in global memory; a producer thread streams them through a ring of
// shared-memory stages with cp.async.bulk (TMA engine) + mbarriers, several tiles ahead, so the consumer warps
// never wait on a global load and never hit a block-wide barrier in the steady state.  A consumer thread keeps a
// fixed 16-byte column position: per row it reads one code byte and one 16-byte piece (a piece that straddles two
// subquantizers takes its upper half with a second code byte and an 8-byte read) and issues one 16-byte store.
// Consecutive lanes own consecutive pieces of a row -> a warp store writes 512 contiguous bytes.  Blocks of the
// column groups of one strip have adjacent block indices: they run concurrently (as a cluster in lockstep when
// the stripes are not whole lines) and the sectors straddling a group boundary are completed in L2.
// ---------------------------------------------------------------------------------------------------------
constexpr int kMaxTileStages = 8;
#ifndef RB_GATHER_UNROLL
#define RB_GATHER_UNROLL 4
#endif
constexpr int kGatherUnroll = RB_GATHER_UNROLL;

// CONS consumer threads + one producer warp; two blocks of 512 + 32 per SM, or one block of 992 + 32 (the
// 1024-thread limit) holding twice the columns (fewer, wider column groups).
template <bool CHECK, int CONS>
__global__ void __launch_bounds__(CONS + 32, CONS == 512 ? 2 : 1)
gather_tile_kernel(const float *__restrict__ quantizers, int k, int dsub, int M, const uint8_t *__restrict__ codes,
                   long long n, float *__restrict__ out, long long ldo, int gw, int n_groups, int W, int tile_rows,
                   int stages, long long tiles_per_strip, long long tile_step, int *__restrict__ err_flag, int lockstep)
{
    using namespace ptx;
    constexpr int kTileConsumers = CONS, kTileThreads = CONS + 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int g = blockIdx.x % n_groups;
    const long long strip = blockIdx.x / n_groups;
    const int col_lo = g * gw;
    const int w = min(gw, M * dsub - col_lo);  // columns of this group (multiple of 4)
    const int w4 = w / 4;                      // 16-byte pieces per row in this group (<= kTileConsumers)

    float *cbt = reinterpret_cast<float *>(smem_raw);               // [k][W]
    uint8_t *ring = reinterpret_cast<uint8_t *>(cbt + (size_t)k * W);  // [stages][tile_rows * M]
    const int stage_bytes = tile_rows * M;                          // multiple of 16
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + (size_t)stages * stage_bytes);
    uint64_t *empty = full + stages;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kTileConsumers / 32);
        }
        fence_mbar_init();
    }
    for (int j = lane; j < w; j += 32) {  // column j of the group: subquantizer m, component t
        const int gc = col_lo + j, m = gc / dsub, t = gc - m * dsub;
        const float *src = quantizers + (size_t)m * k * dsub + t;
        for (int c = warp; c < k; c += kTileThreads / 32) cbt[c * W + j] = __ldg(src + (size_t)c * dsub);
    }
    __syncthreads();

    // strip s walks tiles tile0, tile0 + tile_step, ... (< tile1): contiguous strips (step 1) or interleaved ones
    // (step = number of strips), where the blocks sweep the output together like one linear write stream
    const long long n_tiles = (n + tile_rows - 1) / tile_rows;
    const long long tile0 = tile_step == 1 ? strip * tiles_per_strip : strip;
    const long long tile1 = tile_step == 1 ? min(n_tiles, tile0 + tiles_per_strip) : n_tiles;
    const size_t total_code_bytes = (size_t)n * M;

    if (warp == kTileConsumers / 32) {
        // ===================== producer =====================
        // keeps `stages` tiles in flight; in lockstep mode the whole warp also takes part in the per-tile
        // cluster barrier (every thread of the cluster must arrive)
        int ps = 0;
        uint32_t pph = 1;  // parity of the "empty" phase the next issue waits for
        auto issue = [&](long long tile) {
            mbar_wait(&empty[ps], pph);
            const size_t byte0 = (size_t)tile * stage_bytes;
            const size_t avail = total_code_bytes - byte0;
            const uint32_t bytes = (uint32_t)(avail < (size_t)stage_bytes ? avail : (size_t)stage_bytes);
            const uint32_t bulk = bytes & ~15u;
            uint8_t *dst = ring + (size_t)ps * stage_bytes;
            for (uint32_t b = bulk; b < bytes; b++) dst[b] = codes[byte0 + b];  // ragged end of the matrix
            if (bulk) {
                mbar_arrive_expect_tx(&full[ps], bulk);
                bulk_g2s(dst, codes + byte0, bulk, &full[ps]);
            } else {
                mbar_arrive(&full[ps]);
            }
            if (++ps == stages) {
                ps = 0;
                pph ^= 1;
            }
        };
        if (lane == 0)
            for (long long tile = tile0, i = 0; tile < tile1 && i < stages; tile += tile_step, i++) issue(tile);
        for (long long tile = tile0; tile < tile1; tile += tile_step) {
            __syncwarp();
            if (lockstep) cluster_sync_relaxed();
            if (lane == 0 && tile + stages * tile_step < tile1) issue(tile + stages * tile_step);
        }
        return;
    }

    // ===================== consumers =====================
    const int rows_par = kTileConsumers / w4;
    const int rl = (int)threadIdx.x / w4, c4 = (int)threadIdx.x % w4;
    const bool active = rl < rows_par;
    // the subquantizers of this thread's two 8-byte halves (loop invariant; dsub is even, so a half never straddles)
    const int ma = (col_lo + 4 * c4) / dsub, mb = (col_lo + 4 * c4 + 2) / dsub;
    const bool straddle = ma != mb;
    const float *src = cbt + 4 * c4;
    float *ocol = out + col_lo + 4 * c4;
    bool bad = false;

    int s = 0;
    uint32_t ph = 0;
    for (long long tile = tile0; tile < tile1; tile += tile_step) {
        mbar_wait(&full[s], ph);
        const long long r0 = tile * tile_rows;
        const int rows = (int)min((long long)tile_rows, n - r0);
        if (active) {
            const uint8_t *crow = ring + (size_t)s * stage_bytes + rl * M;
            float *op = ocol + (r0 + rl) * ldo;
            const long long ostep = (long long)rows_par * ldo;
            const int cstep = rows_par * M;
#pragma unroll kGatherUnroll
            for (int row = rl; row < rows; row += rows_par, crow += cstep, op += ostep) {
                unsigned ca = crow[ma];
                if constexpr (CHECK) {
                    bad |= ca >= (unsigned)k;
                    ca = min(ca, (unsigned)(k - 1));
                }
                float4 v = *reinterpret_cast<const float4 *>(src + ca * W);
                if (straddle) {
                    unsigned cb = crow[mb];
                    if constexpr (CHECK) {
                        bad |= cb >= (unsigned)k;
                        cb = min(cb, (unsigned)(k - 1));
                    }
                    const float2 hi = *reinterpret_cast<const float2 *>(src + cb * W + 2);
                    v.z = hi.x;
                    v.w = hi.y;
                }
                *reinterpret_cast<float4 *>(op) = v;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (++s == stages) {
            s = 0;
            ph ^= 1;
        }
        if (lockstep) cluster_sync_relaxed();
    }
    if (bad) atomicExch(err_flag, 1);
}

// Returns true when the tiled kernel was launched (otherwise the caller uses the generic kernel).
bool launch_tiled(const DeviceCodebook &cb, const void *codes, int code_width, size_t n, ptrdiff_t crs, ptrdiff_t ccs,
                  float *out, ptrdiff_t ldo, int *err_flag, cudaStream_t stream, rb_status *status)
{
    const int M = (int)cb.M, k = (int)cb.k, dsub = (int)cb.dsub, d = M * dsub;
    *status = RB_OK;
    if (code_width != 1 || ccs != 1 || crs != (ptrdiff_t)M || (dsub & 1) || (d & 3) || k > 256) return false;
    if ((reinterpret_cast<uintptr_t>(out) & 15) || (ldo & 3) || (reinterpret_cast<uintptr_t>(codes) & 15)) return false;
    if (n < 256) return false;
    struct Shape {
        int cons = 0, gw = 0, W = 0, n_groups = 0, tile_rows = 0, stages = 0;
        size_t smem = 0;
    };
    // development knobs for scripts/gather_time.py sweeps (not part of the ABI): force the block shape, cap the group
    // width, set the ring depth
    static const int env_big = getenv("RB_GATHER_BIG") ? atoi(getenv("RB_GATHER_BIG")) : -1;
    static const int env_w = getenv("RB_GATHER_W") ? atoi(getenv("RB_GATHER_W")) : 0;
    static const int env_stages = getenv("RB_GATHER_STAGES") ? atoi(getenv("RB_GATHER_STAGES")) : 0;
    const int stages = env_stages >= 2 && env_stages <= kMaxTileStages ? env_stages : 3;
    // column groups for a block of `cons` consumers with `smem_budget` bytes of shared memory, of which at most
    // `cb_budget` hold the codebook slice; n_groups == 0: no fit
    auto shape_for = [&](int cons, size_t smem_budget, size_t cb_budget) {
        Shape sh;
        sh.cons = cons;
        sh.stages = stages;
        int wmax = (int)(cb_budget / ((size_t)k * sizeof(float))) / 32 * 32;  // widest pitch that fits
        if (wmax > cons * 4 / 32 * 32) wmax = cons * 4 / 32 * 32;              // one 16-byte piece per thread
        if (env_w >= 32 && env_w < wmax) wmax = env_w / 32 * 32;
        if (wmax < 32) return sh;
        int n_groups = (int)ceil_div(d, wmax);
        int gw = (int)ceil_div(d, n_groups);
        gw += (4 - gw % 4) % 4;
        // whole 128-byte lines per stripe when the rows allow it
        if (d % 32 == 0 && (gw + 31) / 32 * 32 <= wmax && ceil_div(d, (gw + 31) / 32 * 32) == (size_t)n_groups)
            gw = (gw + 31) / 32 * 32;
        n_groups = (int)ceil_div(d, gw);
        const int W = (gw + 31) / 32 * 32;
        const size_t cb_bytes = (size_t)k * W * sizeof(float);
        const size_t bar_bytes = 2 * stages * sizeof(uint64_t);
        if (cb_bytes + bar_bytes + (size_t)stages * 16 * M > smem_budget) return sh;
        int tile_rows = (int)((smem_budget - cb_bytes - bar_bytes) / stages / M);
        if (tile_rows > 256) tile_rows = 256;
        // whole passes of the consumers over a tile (rows_par rows per pass), stage size a multiple of 16 bytes
        const int rows_par = cons / (gw / 4);
        int best = 0;
        for (int t = tile_rows; t >= 8 && t > tile_rows - 4 * rows_par; t--)
            if (((size_t)t * M) % 16 == 0 && (best == 0 || (t % rows_par == 0))) {
                if (best == 0 || t % rows_par == 0) best = t;
                if (t % rows_par == 0) break;
            }
        if (best == 0) return sh;
        tile_rows = best;
        sh.gw = gw;
        sh.W = W;
        sh.n_groups = n_groups;
        sh.tile_rows = tile_rows;
        sh.smem = cb_bytes + (size_t)stages * tile_rows * M + bar_bytes;
        return sh;
    };
    auto aligned = [&](const Shape &s) {
        return (reinterpret_cast<uintptr_t>(out) % 128 == 0) && (ldo % 32 == 0) && (s.gw % 32 == 0);
    };
    Shape sh = shape_for(512, 113 * 1024, 100 * 1024);         // two blocks per SM
    const Shape big = shape_for(992, 225 * 1024, 200 * 1024);  // one block per SM
    // the big block pays off when it removes column groups whose stripes are not whole 128-byte lines
    if (big.n_groups > 0 && (sh.n_groups == 0 || env_big == 1 || (big.n_groups < sh.n_groups && !aligned(sh)))) {
        if (env_big != 0 || sh.n_groups == 0) sh = big;
    }
    if (sh.n_groups == 0) return false;
    const int n_groups = sh.n_groups, tile_rows = sh.tile_rows, blocks_per_sm = sh.cons == 512 ? 2 : 1;
    const size_t smem = sh.smem;

    const bool check = k < 256;
    auto kern = sh.cons == 512 ? (check ? gather_tile_kernel<true, 512> : gather_tile_kernel<false, 512>)
                               : (check ? gather_tile_kernel<true, 992> : gather_tile_kernel<false, 992>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        *status = RB_ERR_CUDA;
        return true;
    }
    // The column groups of one strip form a thread-block cluster and advance tile by tile in lockstep, so the
    // stripes of a row region reach L2 within microseconds of each other and leave it as complete lines.  Not
    // needed when every stripe segment is a whole number of 128-byte lines anyway.
    const int lockstep = (n_groups > 1 && n_groups <= 8 && !aligned(sh)) ? 1 : 0;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3((unsigned)sh.cons + 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = lockstep ? (unsigned)n_groups : 1u;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;

    // one resident wave: as many strips as clusters (or blocks) fit on the device at once
    const long long n_tiles = (long long)ceil_div(n, (size_t)tile_rows);
    long long strips = ((long long)blocks_per_sm * sm_count()) / n_groups;
    if (lockstep) {
        int max_clusters = 0;
        cfg.gridDim = dim3((unsigned)(strips * n_groups));
        if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) == cudaSuccess && max_clusters > 0 &&
            max_clusters < strips)
            strips = max_clusters;
        (void)cudaGetLastError();
    }
    if (strips < 1) strips = 1;
    if (strips > n_tiles) strips = n_tiles;
    const long long tiles_per_strip = (n_tiles + strips - 1) / strips;
    strips = (n_tiles + tiles_per_strip - 1) / tiles_per_strip;
    const unsigned grid = (unsigned)(strips * n_groups);
    cfg.gridDim = dim3(grid);
    static const int env_interleave = getenv("RB_GATHER_INTERLEAVE") ? atoi(getenv("RB_GATHER_INTERLEAVE")) : 1;
    const long long tile_step = env_interleave ? strips : 1;
    e = cudaLaunchKernelEx(&cfg, kern, cb.quantizers, k, dsub, M, reinterpret_cast<const uint8_t *>(codes),
                           (long long)n, out, (long long)ldo, sh.gw, n_groups, sh.W, tile_rows, sh.stages, tiles_per_strip, tile_step,
                           err_flag, lockstep);
    if (e != cudaSuccess && lockstep) {
        // the cluster shape could not be placed (partitioned GPU, ...): same kernel without the lockstep barrier
        (void)cudaGetLastError();
        attr[0].val.clusterDim.x = 1;
        e = cudaLaunchKernelEx(&cfg, kern, cb.quantizers, k, dsub, M, reinterpret_cast<const uint8_t *>(codes),
                               (long long)n, out, (long long)ldo, sh.gw, n_groups, sh.W, tile_rows, sh.stages, tiles_per_strip, tile_step,
                               err_flag, 0);
    }
    if (e != cudaSuccess) {
        set_error("cudaLaunchKernelEx failed: %s (%s:%d)", cudaGetErrorString(e), __FILE__, __LINE__);
        *status = RB_ERR_CUDA;
        return true;
    }
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
    size_t strips = ceil_div((size_t)sm_count() * 4, (size_t)n_groups);
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
