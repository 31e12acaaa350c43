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
// Two accumulate paths produce the same packed layout:
//   ordered (default, k <= 1024): reproduces the reference's summation ORDER.  kmeans.rs:185-189 adds the rows
//     of a cluster sequentially in increasing row order in f32; k-means is chaotic (one last-bit difference in
//     a centroid flips a near-tie assignment, which moves two centroids by 1/count), so only the same order
//     keeps trained centroids equal to the oracle's.  Implemented as a stable counting sort of the row indices
//     by code (per-chunk histograms -> scan -> stable warp-synchronous scatter) followed by one sequential
//     f32 chain per (subquantizer, cluster, component) — M*k*dsub independent chains, 8 gathers in flight each.
//     Result: sums, hence trained centroids, are BIT-IDENTICAL to the oracle on one GPU.
//   atomic: shared-memory atomics, summation order unspecified (FP32 order-of-summation error ~1e-7 relative);
//     used for k > 1024.
// Across ranks (data-parallel k-means) the all-reduce adds per-rank partial sums, which is not the
// single sequential chain: that mode is tolerance-level.  The `init` argument of the ordered path continues
// chains begun on the preceding rank, which keeps multi-GPU training bit-exact (dist.py mode="chained").
// Counts follow the reference's f32 `+= 1.0` (exact to 2^24 per cluster, kmeans.rs:188).
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace rb {

namespace {

constexpr int kAccThreads = 256;

template <bool SMEM_ACC, typename CodeT>
__global__ void __launch_bounds__(kAccThreads)
accumulate_kernel(const float *__restrict__ x, long long n, long long ldx, const CodeT *__restrict__ codes,
                  long long code_pitch, int M, int k, int dsub, float *__restrict__ packed, long long rows_per_block)
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
                const int code = (int)codes[(long long)m * code_pitch + row];
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


// ---- ordered path ---------------------------------------------------------------------------------------
constexpr int kSortWarps = 4;  // warps per block; each warp owns one (row chunk, subquantizer) pair

// Pass 1: per (chunk, m) histogram of codes.  cnt layout [chunk][m][k].
template <typename CodeT>
__global__ void __launch_bounds__(kSortWarps * 32)
sort_hist_kernel(const CodeT *__restrict__ codes, long long code_pitch, long long n, int M, int k,
                 long long rows_per_chunk, int n_chunks, unsigned *__restrict__ cnt)
{
    extern __shared__ unsigned sh[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long pair = (long long)blockIdx.x * kSortWarps + warp;
    unsigned *h = sh + (size_t)warp * k;
    for (int i = lane; i < k; i += 32) h[i] = 0;
    __syncwarp();
    if (pair < (long long)n_chunks * M) {
        const int chunk = (int)(pair / M), m = (int)(pair % M);
        const long long r0 = (long long)chunk * rows_per_chunk, r1 = min(n, r0 + rows_per_chunk);
        for (long long row = r0 + lane; row < r1; row += 32) atomicAdd(h + (unsigned)codes[(long long)m * code_pitch + row], 1u);
        __syncwarp();
        unsigned *dst = cnt + (size_t)pair * k;
        for (int i = lane; i < k; i += 32) dst[i] = h[i];
    }
}

// Pass 2: for every (m, j): exclusive scan over chunks (in place: cnt becomes the chunk's start offset inside
// the cluster's list) and the cluster total; then per m an exclusive scan over j gives the list bases.
__global__ void sort_scan_kernel(unsigned *__restrict__ cnt, int M, int k, int n_chunks, unsigned *__restrict__ total,
                                 unsigned *__restrict__ base)
{
    const int m = blockIdx.x;
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        unsigned run = 0;
        for (int c = 0; c < n_chunks; c++) {
            unsigned *p = cnt + ((size_t)c * M + m) * k + j;
            const unsigned v = *p;
            *p = run;
            run += v;
        }
        total[(size_t)m * k + j] = run;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned run = 0;
        for (int j = 0; j < k; j++) {
            base[(size_t)m * k + j] = run;
            run += total[(size_t)m * k + j];
        }
    }
}

// Pass 3: stable scatter of row indices.  list layout [m][n]; cluster j of subquantizer m occupies
// list[m][base[m][j] .. + total[m][j]) in increasing row order.
template <typename CodeT>
__global__ void __launch_bounds__(kSortWarps * 32)
sort_scatter_kernel(const CodeT *__restrict__ codes, long long code_pitch, long long n, int M, int k,
                    long long rows_per_chunk, int n_chunks, const unsigned *__restrict__ off,
                    const unsigned *__restrict__ base, unsigned *__restrict__ list)
{
    extern __shared__ unsigned sh[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long pair = (long long)blockIdx.x * kSortWarps + warp;
    if (pair >= (long long)n_chunks * M) return;
    const int chunk = (int)(pair / M), m = (int)(pair % M);
    unsigned *pos = sh + (size_t)warp * k;  // next free slot of every cluster, for this chunk
    for (int i = lane; i < k; i += 32) pos[i] = base[(size_t)m * k + i] + off[(size_t)pair * k + i];
    __syncwarp();
    const long long r0 = (long long)chunk * rows_per_chunk, r1 = min(n, r0 + rows_per_chunk);
    unsigned *lm = list + (size_t)m * n;
    for (long long rb = r0; rb < r1; rb += 32) {
        const long long row = rb + lane;
        const bool live = row < r1;
        const unsigned code = live ? (unsigned)codes[(long long)m * code_pitch + row] : 0xffffffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, code);  // lanes (= consecutive rows) with my code
        if (live) {
            const unsigned rank = __popc(peers & ((1u << lane) - 1u));
            lm[pos[code] + rank] = (unsigned)row;
        }
        __syncwarp();
        if (live && (peers >> lane) == 1u) pos[code] += __popc(peers);  // highest lane of the group advances
        __syncwarp();
    }
}

// Pass 4: one sequential f32 chain per (m, j, t) in row order (kmeans.rs:185-189), 8 gathers in flight.
__global__ void __launch_bounds__(256)
ordered_sum_kernel(const float *__restrict__ x, long long ldx, long long n, int M, int k, int dsub,
                   const unsigned *__restrict__ list, const unsigned *__restrict__ total,
                   const unsigned *__restrict__ base, float *__restrict__ packed)
{
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long chains = (long long)M * k * dsub;
    double sq = 0.0;
    int m = 0;
    if (tid < chains) {
        const int t = (int)(tid % dsub);
        const long long mj = tid / dsub;
        m = (int)(mj / k);
        const unsigned cnt = total[mj];
        const unsigned *l = list + (size_t)m * n + base[mj];
        const float *xc = x + (long long)m * dsub + t;
        float acc = 0.f;
        unsigned i = 0;
        for (; i + 8 <= cnt; i += 8) {
            unsigned r[8];
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) r[u] = __ldg(l + i + u);
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = __ldg(xc + (long long)r[u] * ldx);
#pragma unroll
            for (int u = 0; u < 8; u++) {
                acc = __fadd_rn(acc, v[u]);
                sq += (double)v[u] * (double)v[u];
            }
        }
        for (; i < cnt; i++) {
            const float v = __ldg(xc + (long long)__ldg(l + i) * ldx);
            acc = __fadd_rn(acc, v);
            sq += (double)v * (double)v;
        }
        packed[tid] = acc;
        if (t == 0) packed[chains + mj] = (float)cnt;
    }
    // squared-norm partials: threads of one warp may straddle two subquantizers only when k*dsub % 32 != 0;
    // use per-thread atomics in that (rare) case, a warp reduction otherwise.
    float *gsq = packed + chains + (long long)M * k;
    const bool uniform = ((long long)k * dsub) % 32 == 0;
    if (uniform) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
        if ((threadIdx.x & 31) == 0 && tid < chains) atomicAdd(gsq + m, (float)sq);
    } else if (tid < chains) {
        atomicAdd(gsq + m, (float)sq);
    }
}


// ---- ordered path, fast variant (u8 codes, k <= 256, dsub <= 32) --------------------------------------------
// Same result as the generic ordered path, organised for the memory system:
//   sort_local_kernel  one block per (chunk of 4 096 - 65 536 rows, subquantizer): stable counting sort of the chunk's row
//     offsets by code entirely in shared memory (histogram with shared atomics; stable placement with warp-sequential
//     steps whose same-code groups come from 8 ballots, one per code bit), written out as one contiguous, fully
//     coalesced u16 segment plus the k+1 cluster boundaries of the chunk;
//   ordered_chain_kernel  launched once per chunk; lanes = (chain, component): a warp carries 32 / dsub clusters, each
//     lane walks its own cluster's rows in order with 8 gathers in flight and does the reference's sequential f32
//     adds (kmeans.rs:185-189) without any cross-lane traffic; the running sums live in `packed` between launches.  Because every chain is in
//     the same chunk at the same time, the chunk's rows (< 96 MB, inside L2) are fetched from HBM once although
//     each (row, subquantizer) piece is gathered by a different warp.  `init` (or nullptr) continues chains started
//     on another rank: data-parallel training can pass the running sums from rank to rank and stay bit-identical.
constexpr int kLocalThreads = 1024;  // one block per SM (shared memory): 32 warps hide the serial placement steps
constexpr int kLocalWarps = kLocalThreads / 32;

// rows per chunk: a power of two in [4096, 65536] (u16 offsets) whose x rows (4*d bytes each) stay below ~96 MB of L2
#ifndef RB_CHUNK_MB
#define RB_CHUNK_MB 96
#endif
int chunk_rows_for(size_t d)
{
    size_t rows = ((size_t)RB_CHUNK_MB << 20) / (4 * (d ? d : 1));
    int p = 4096;
    while (p < 65536 && (size_t)(2 * p) <= rows) p *= 2;
    return p;
}

// lanes holding the same 8-bit code as this lane (what __match_any_sync computes, which is far slower)
__device__ __forceinline__ unsigned same_code_lanes(unsigned code, bool live)
{
    unsigned peers = __ballot_sync(0xffffffffu, live);
#pragma unroll
    for (int b = 0; b < 8; b++) {
        const unsigned bit = (code >> b) & 1u;
        const unsigned v = __ballot_sync(0xffffffffu, bit);
        peers &= v ^ (bit - 1u);  // one three-input logic op: lanes whose bit b equals mine
    }
    return peers;
}

__global__ void __launch_bounds__(kLocalThreads)
sort_local_kernel(const uint8_t *__restrict__ codes, long long code_pitch, long long n, int M, int k, int n_chunks,
                  int kChunkRows, uint16_t *__restrict__ list, uint32_t *__restrict__ lstart)
{
    const int kRowsPerWarp = kChunkRows / kLocalWarps;
    extern __shared__ __align__(16) unsigned char sm_raw[];
    uint8_t *sc = sm_raw;                                                   // [kChunkRows] codes of the chunk
    uint16_t *sorted = reinterpret_cast<uint16_t *>(sm_raw + kChunkRows);   // [kChunkRows] row offsets by code
    uint32_t *wcnt = reinterpret_cast<uint32_t *>(sorted + kChunkRows);     // [kLocalWarps][k]
    uint32_t *cstart = wcnt + kLocalWarps * k;                              // [k + 1]
    const int chunk = blockIdx.x, m = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long r0 = (long long)chunk * kChunkRows;
    const int rows = (int)min((long long)kChunkRows, n - r0);
    const uint8_t *col = codes + (long long)m * code_pitch + r0;  // 16-byte aligned (pitch and chunk are multiples of 16)

    for (int i = threadIdx.x * 16; i < rows; i += kLocalThreads * 16)
        *reinterpret_cast<uint4 *>(sc + i) = __ldg(reinterpret_cast<const uint4 *>(col + i));  // pitch padding covers the tail
    for (int i = threadIdx.x; i < kLocalWarps * k; i += kLocalThreads) wcnt[i] = 0;
    __syncthreads();

    // pass 1: per-warp histograms (warp w owns rows [w * kRowsPerWarp, (w + 1) * kRowsPerWarp))
    uint32_t *mine = wcnt + warp * k;
    const int w0 = warp * kRowsPerWarp;
    for (int s = 0; s < kRowsPerWarp; s += 32) {
        const int i = w0 + s + lane;
        if (i < rows) atomicAdd(mine + sc[i], 1u);
    }
    __syncthreads();
    // exclusive scan over warps per cluster, then over clusters
    for (int j = threadIdx.x; j < k; j += kLocalThreads) {
        unsigned run = 0;
        for (int w = 0; w < kLocalWarps; w++) {
            const unsigned c = wcnt[w * k + j];
            wcnt[w * k + j] = run;
            run += c;
        }
        cstart[j + 1] = run;  // cluster total for now
    }
    __syncthreads();
    if (warp == 0) {  // k <= 256 totals: a few per lane, then a warp scan
        const int per = (k + 31) / 32;
        unsigned local = 0;
        for (int q = 0; q < per; q++) {
            const int j = lane * per + q;
            if (j < k) local += cstart[j + 1];
        }
        unsigned incl = local;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned o = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += o;
        }
        unsigned run = incl - local;
        for (int q = 0; q < per; q++) {
            const int j = lane * per + q;
            if (j < k) {
                run += cstart[j + 1];
                cstart[j + 1] = run;  // END of cluster j == start of cluster j + 1
            }
        }
        if (lane == 0) cstart[0] = 0;
    }
    __syncthreads();
    // pass 2: stable placement (rows of a warp in order, lanes of a step in order)
    for (int s = 0; s < kRowsPerWarp; s += 32) {
        const int i = w0 + s + lane;
        const bool live = i < rows;
        const unsigned code = live ? sc[i] : 0u;
        const unsigned peers = same_code_lanes(code, live);
        if (live) {
            const unsigned rank = __popc(peers & ((1u << lane) - 1u));
            sorted[cstart[code] + mine[code] + rank] = (uint16_t)i;
        }
        __syncwarp();
        if (live && (peers >> lane) == 1u) mine[code] += __popc(peers);  // highest lane of the group advances
        __syncwarp();
    }
    __syncthreads();
    uint16_t *dst = list + ((size_t)m * n_chunks + chunk) * kChunkRows;
    for (int i = threadIdx.x * 8; i < rows; i += kLocalThreads * 8)
        *reinterpret_cast<uint4 *>(dst + i) = *reinterpret_cast<const uint4 *>(sorted + i);
    uint32_t *ls = lstart + ((size_t)chunk * M + m) * (k + 1);
    for (int j = threadIdx.x; j <= k; j += kLocalThreads) ls[j] = cstart[j];
}

// read-only load that asks L2 to bring in the whole 128-byte line: the line's other three 32-byte pieces belong to
// neighbouring subquantizers of the same row and are gathered by other warps during the same launch
// programmatic dependent launch: wait for the previous launch of the stream / let the next one start early
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float ldg_l2_128(const float *p)
{
    float v;
    asm volatile("ld.global.nc.L2::128B.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

template <int DSUB, int UN>
__global__ void __launch_bounds__(256)
ordered_chain_kernel(const float *__restrict__ x, long long ldx, int M, int k, int n_chunks, int chunk, int kChunkRows,
                     const uint16_t *__restrict__ list, const uint32_t *__restrict__ lstart,
                     const float *__restrict__ init, float *__restrict__ packed)
{
    // lanes = (chain c, component t): a warp carries 32 / DSUB independent chains, i.e. clusters; every lane walks
    // its own cluster's rows of this chunk in order, so the reference's sequential adds need no cross-lane traffic
    constexpr int CH = 32 / DSUB;  // chains per warp
    // UN = rows in flight per chain: 8 when the GPU carries many chains (a whole training matrix), 32 when it carries
    // few (the subquantizer shard of a multi-GPU run: a launch is then bound by the latency of its batches)
    const int lane = threadIdx.x & 31;
    const int c = lane / DSUB, t = lane % DSUB;
    const long long total_mj = (long long)M * k;
    const long long mj = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * CH + c;
    const bool on = c < CH && mj < total_mj;
    const long long chains = total_mj * DSUB;
    const long long mjc = on ? mj : 0;
    const int m = (int)(mjc / k), j = (int)(mjc % k);
    // Launched with programmatic stream serialization after the first chunk: everything up to griddep_wait() runs
    // while the previous chunk's launch is still draining -- the cluster boundaries, the first row offsets and the
    // first batch of gathers depend only on the sort and on x, which were complete before the first chain launch
    // passed its own wait.  Only the running sums (prev) need the previous launch.
    if (chunk == 0) griddep_wait();
    griddep_launch_dependents();
    unsigned s = 0, e = 0;
    if (on) {
        const uint32_t *ls = lstart + ((size_t)chunk * M + m) * (k + 1) + j;
        s = __ldg(ls);
        e = __ldg(ls + 1);
    }
    // byte offsets in 32 bits where possible: a row offset inside the chunk is < 65 536 and the row pitch in bytes fits
    // 32 bits, so one IMAD.WIDE.U32 per gather forms the address
    const uint16_t *sp = list + ((size_t)m * n_chunks + chunk) * kChunkRows + s;
    const char *xc = reinterpret_cast<const char *>(x + (long long)chunk * kChunkRows * ldx + (long long)m * DSUB + t);
    const unsigned pitch_b = (unsigned)ldx * 4u;
    float sqf = 0.f;  // this lane's share of sum ||x||^2 in this chunk (<= a few hundred terms), widened once below
    int rem = (int)(e - s);
    const uint16_t *seg0 = list + ((size_t)m * n_chunks + chunk) * kChunkRows;  // always readable
    // Loads are unconditional (indices clamped to the last valid row, finished chains re-read the segment start); only
    // the adds are predicated.  The row offsets of batch b + 1 are fetched while batch b's gathers are in flight, so a
    // batch costs one memory round trip, not two.
    unsigned off[UN];
    float v[UN];
    auto fetch_offsets = [&](int left) {  // offsets of the batch that starts at sp with `left` rows still to add
        const uint16_t *spu = left > 0 ? sp : seg0;
        const int last = max(min(left, UN), 1) - 1;
#pragma unroll
        for (int u = 0; u < UN; u++) off[u] = (unsigned)__ldg(spu + min(u, last));
    };
    auto gather = [&]() {
#pragma unroll
        for (int u = 0; u < UN; u++) v[u] = ldg_l2_128(reinterpret_cast<const float *>(xc + (unsigned long long)off[u] * (unsigned long long)pitch_b));
    };
    fetch_offsets(rem);
    gather();
    sp += UN;
    int rem_next = rem - UN;
    fetch_offsets(rem_next);
    if (chunk != 0) griddep_wait();
    // the chain so far: other ranks' rows (init) before this rank's first chunk, else what the last launch left
    const float *prev = chunk == 0 ? init : packed;
    float acc = (on && prev) ? prev[mjc * DSUB + t] : 0.f;
    double sq = 0.0;
    while (__any_sync(0xffffffffu, rem > 0)) {
#pragma unroll
        for (int u = 0; u < UN; u++) {
            if (u < rem) {
                acc = __fadd_rn(acc, v[u]);  // kmeans.rs:185-189: one rounded add per row, in row order
                sqf = fmaf(v[u], v[u], sqf);
            }
        }
        rem = rem_next;
        gather();
        sp += UN;
        rem_next = rem - UN;
        fetch_offsets(rem_next);
    }
    sq = (double)sqf;
    if (on) {
        packed[mj * DSUB + t] = acc;
        if (t == 0) packed[chains + mj] = (prev ? prev[chains + mj] : 0.f) + (float)(e - s);
    }
    // sum of squares: reduce over the chain's DSUB lanes (they are consecutive), one atomic per chain
#pragma unroll
    for (int off2 = 1; off2 < DSUB; off2 <<= 1) {
        const double o = __shfl_down_sync(0xffffffffu, sq, off2);
        if (t + off2 < DSUB) sq += o;
    }
    if (on && t == 0 && sq != 0.0) atomicAdd(packed + chains + total_mj + m, (float)sq);
}

bool fast_ordered_supported(size_t n, size_t k, size_t dsub)
{
    if (k > 256 || n >= ((size_t)1 << 40)) return false;
    switch (dsub) {
    case 1: case 2: case 4: case 8: case 16: case 32: case 3: case 5: case 6: case 10: case 12: case 15: case 20: case 30:
        return true;
    default: return false;
    }
}

rb_status launch_ordered_fast(const float *x, size_t n, ptrdiff_t ldx, const uint8_t *codes, size_t code_pitch, size_t M,
                              size_t k, size_t dsub, const float *init, float *packed, cudaStream_t stream)
{
    const int kChunkRows = chunk_rows_for(M * dsub);
    const size_t n_chunks = ceil_div(n, (size_t)kChunkRows);
    uint16_t *list = nullptr;
    uint32_t *lstart = nullptr;
    RB_CUDA_TRY(pool_malloc((void **)&list, M * n_chunks * kChunkRows * sizeof(uint16_t), stream));
    RB_CUDA_TRY(pool_malloc((void **)&lstart, n_chunks * M * (k + 1) * sizeof(uint32_t), stream));
    rb_status st = RB_OK;
    auto body = [&]() -> rb_status {
        const size_t smem = kChunkRows + (size_t)kChunkRows * 2 + (size_t)kLocalWarps * k * 4 + (k + 1) * 4;
        RB_CUDA_TRY(cudaFuncSetAttribute(sort_local_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sort_local_kernel<<<dim3((unsigned)n_chunks, (unsigned)M), kLocalThreads, smem, stream>>>(
            codes, (long long)code_pitch, (long long)n, (int)M, (int)k, (int)n_chunks, kChunkRows, list, lstart);
        RB_LAUNCH_CHECK();
        if (init) {  // the sum of squared norms continues as well
            RB_CUDA_TRY(cudaMemcpyAsync(packed + M * k * dsub + M * k, init + M * k * dsub + M * k, M * sizeof(float),
                                        cudaMemcpyDeviceToDevice, stream));
        }
        const unsigned blocks = (unsigned)ceil_div(ceil_div(M * k, 32 / dsub), 8);
        const bool deep = M * k * dsub <= (size_t)sm_count() * 512;  // fewer than ~16 chain warps per SM
        // chunk 0 is an ordinary launch (it needs the sort); the others may start while their predecessor drains
        cudaLaunchAttribute pdl[1];
        pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        pdl[0].val.programmaticStreamSerializationAllowed = 1;
        for (size_t c = 0; c < n_chunks; c++) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(blocks);
            cfg.blockDim = dim3(256);
            cfg.dynamicSmemBytes = 0;
            cfg.stream = stream;
            cfg.attrs = pdl;
            cfg.numAttrs = c > 0 ? 1 : 0;
#define RB_CHAIN(D)                                                                                                  \
    case D:                                                                                                          \
        if (deep)                                                                                                    \
            RB_CUDA_TRY(cudaLaunchKernelEx(&cfg, ordered_chain_kernel<D, 32>, x, (long long)ldx, (int)M, (int)k,     \
                                           (int)n_chunks, (int)c, kChunkRows, (const uint16_t *)list,                \
                                           (const uint32_t *)lstart, init, packed));                                 \
        else                                                                                                         \
            RB_CUDA_TRY(cudaLaunchKernelEx(&cfg, ordered_chain_kernel<D, 8>, x, (long long)ldx, (int)M, (int)k,      \
                                           (int)n_chunks, (int)c, kChunkRows, (const uint16_t *)list,                \
                                           (const uint32_t *)lstart, init, packed));                                 \
        break;
            switch (dsub) {
                RB_CHAIN(1) RB_CHAIN(2) RB_CHAIN(3) RB_CHAIN(4) RB_CHAIN(5) RB_CHAIN(6) RB_CHAIN(8) RB_CHAIN(10)
                RB_CHAIN(12) RB_CHAIN(15) RB_CHAIN(16) RB_CHAIN(20) RB_CHAIN(30) RB_CHAIN(32)
            default: break;
            }
#undef RB_CHAIN
            RB_LAUNCH_CHECK();
        }
        return RB_OK;
    };
    st = body();
    cudaFreeAsync(list, stream);
    cudaFreeAsync(lstart, stream);
    return st;
}

__global__ void __launch_bounds__(256)
finalize_kernel(const float *__restrict__ packed, int M, int k, int dsub, double inv_len, float *__restrict__ centroids,
                float *__restrict__ loss, const double *__restrict__ sumsq64)
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
        // sum ||x||^2: the deterministic FP64 value when the caller has one (training loops: x never changes, so it
        // is computed once per run), else the packed buffer's float slot (accumulated with atomics)
        double s = sumsq64 != nullptr ? sumsq64[m] : (double)sumsq;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += red[w];
        if (s < 0.0) s = 0.0;
        loss[m] = (float)(s * inv_len);  // kmeans.rs:359: sse / (n * dsub)
    }
}

template <typename CodeT>
rb_status launch_acc_t(const float *x, size_t n, ptrdiff_t ldx, const CodeT *codes, size_t code_pitch, size_t M, size_t k,
                       size_t dsub, float *packed, cudaStream_t stream)
{
    const size_t pad = dsub | 1;
    const size_t smem = k * pad * sizeof(float) + k * sizeof(int);
    const bool smem_acc = smem <= 160 * 1024;
    // ~4 blocks per SM across all subquantizers, at least 1024 rows per block
    size_t chunks = ceil_div((size_t)sm_count() * 4, M);
    size_t rows_per_block = ceil_div(n, chunks);
    if (rows_per_block < 1024) rows_per_block = 1024;
    chunks = ceil_div(n, rows_per_block);
    dim3 grid((unsigned)chunks, (unsigned)M);
    if (smem_acc) {
        auto kern = accumulate_kernel<true, CodeT>;
        if (smem > 48 * 1024)
            RB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kAccThreads, smem, stream>>>(x, (long long)n, (long long)ldx, codes, (long long)code_pitch, (int)M,
                                                 (int)k, (int)dsub, packed, (long long)rows_per_block);
    } else {
        accumulate_kernel<false, CodeT><<<grid, kAccThreads, 0, stream>>>(
            x, (long long)n, (long long)ldx, codes, (long long)code_pitch, (int)M, (int)k, (int)dsub, packed,
            (long long)rows_per_block);
    }
    RB_LAUNCH_CHECK();
    return RB_OK;
}


template <typename CodeT>
rb_status launch_ordered_t(const float *x, size_t n, ptrdiff_t ldx, const CodeT *codes, size_t code_pitch, size_t M, size_t k,
                           size_t dsub, float *packed, cudaStream_t stream)
{
    // chunks: about 2 waves of warps over the GPU, at least 256 rows per chunk
    size_t n_chunks = ceil_div((size_t)sm_count() * 64, M);
    size_t rows_per_chunk = ceil_div(n, n_chunks);
    if (rows_per_chunk < 256) rows_per_chunk = 256;
    n_chunks = ceil_div(n, rows_per_chunk);
    const size_t pairs = n_chunks * M;
    unsigned *cnt = nullptr, *total = nullptr, *base = nullptr, *list = nullptr;
    RB_CUDA_TRY(pool_malloc((void **)&cnt, pairs * k * sizeof(unsigned), stream));
    RB_CUDA_TRY(pool_malloc((void **)&total, M * k * sizeof(unsigned), stream));
    RB_CUDA_TRY(pool_malloc((void **)&base, M * k * sizeof(unsigned), stream));
    RB_CUDA_TRY(pool_malloc((void **)&list, M * n * sizeof(unsigned), stream));
    const size_t smem = (size_t)kSortWarps * k * sizeof(unsigned);
    const unsigned blocks = (unsigned)ceil_div(pairs, kSortWarps);
    rb_status st = RB_OK;
    auto body = [&]() -> rb_status {
        sort_hist_kernel<CodeT><<<blocks, kSortWarps * 32, smem, stream>>>(
            codes, (long long)code_pitch, (long long)n, (int)M, (int)k, (long long)rows_per_chunk, (int)n_chunks, cnt);
        RB_LAUNCH_CHECK();
        sort_scan_kernel<<<(unsigned)M, 256, 0, stream>>>(cnt, (int)M, (int)k, (int)n_chunks, total, base);
        RB_LAUNCH_CHECK();
        sort_scatter_kernel<CodeT><<<blocks, kSortWarps * 32, smem, stream>>>(
            codes, (long long)code_pitch, (long long)n, (int)M, (int)k, (long long)rows_per_chunk, (int)n_chunks, cnt, base,
            list);
        RB_LAUNCH_CHECK();
        const size_t chains = M * k * dsub;
        ordered_sum_kernel<<<(unsigned)ceil_div(chains, 256), 256, 0, stream>>>(
            x, (long long)ldx, (long long)n, (int)M, (int)k, (int)dsub, list, total, base, packed);
        RB_LAUNCH_CHECK();
        return RB_OK;
    };
    st = body();
    cudaFreeAsync(cnt, stream);
    cudaFreeAsync(total, stream);
    cudaFreeAsync(base, stream);
    cudaFreeAsync(list, stream);
    return st;
}

}  // namespace

rb_status launch_kmeans_accumulate(const float *x, size_t n, ptrdiff_t ldx, const uint8_t *codes8,
                                   const uint32_t *codes32, size_t code_pitch, size_t M, size_t k, size_t dsub,
                                   const float *init, float *packed, int ordered, cudaStream_t stream)
{
    const size_t len = rb_kmeans_packed_len(M, k, dsub);
    if (init && !(ordered && codes8 && fast_ordered_supported(n, k, dsub) && ldx > 0 && ldx < ((ptrdiff_t)1 << 29))) {
        set_error("continuing chains (init) needs the ordered update with u8 codes (k <= 256, dsub <= 32)");
        return RB_ERR_UNSUPPORTED;
    }
    if (n == 0) {
        if (init) RB_CUDA_TRY(cudaMemcpyAsync(packed, init, len * sizeof(float), cudaMemcpyDeviceToDevice, stream));
        else RB_CUDA_TRY(cudaMemsetAsync(packed, 0, len * sizeof(float), stream));
        return RB_OK;
    }
    RB_CUDA_TRY(cudaMemsetAsync(packed, 0, len * sizeof(float), stream));
    if (ordered && codes8 && fast_ordered_supported(n, k, dsub) && ldx > 0 && ldx < ((ptrdiff_t)1 << 29))
        return launch_ordered_fast(x, n, ldx, codes8, code_pitch, M, k, dsub, init, packed, stream);
    if (ordered && k <= 1024 && n < ((size_t)1 << 32)) {
        if (codes8) return launch_ordered_t<uint8_t>(x, n, ldx, codes8, code_pitch, M, k, dsub, packed, stream);
        return launch_ordered_t<uint32_t>(x, n, ldx, codes32, code_pitch, M, k, dsub, packed, stream);
    }
    if (codes8) return launch_acc_t<uint8_t>(x, n, ldx, codes8, code_pitch, M, k, dsub, packed, stream);
    return launch_acc_t<uint32_t>(x, n, ldx, codes32, code_pitch, M, k, dsub, packed, stream);
}

rb_status launch_kmeans_finalize(const float *packed, size_t M, size_t k, size_t dsub, uint64_t n_total,
                                 float *centroids, float *loss, cudaStream_t stream, const double *sumsq64)
{
    const double len = (double)n_total * (double)dsub;
    finalize_kernel<<<(unsigned)M, 256, 0, stream>>>(packed, (int)M, (int)k, (int)dsub, len > 0 ? 1.0 / len : 0.0,
                                                     centroids, loss, sumsq64);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

namespace {

// sum over rows of ||x_m||^2 per subquantizer in FP64 with a FIXED summation order (fixed row ranges per block, fixed
// strides per thread, tree reduction, partials added in block order): run-to-run identical, unlike float atomics.
// sum ||x_m||^2 over all rows per subquantizer, in FP64 with a fixed order (the losses that rank training attempts must
// not depend on the run): a thread owns one column of a row split (coalesced row reads, eight in flight), then one
// warp per subquantizer adds the splits and the subquantizer's columns in a fixed pattern.
__global__ void __launch_bounds__(256)
sumsq_partial_kernel(const float *__restrict__ x, long long n, long long ldx, int d, long long rows_per_split,
                     double *__restrict__ partial)
{
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
        for (int e = 0; e < 8; e++) acc += (double)v[e] * (double)v[e];
    }
    for (; r < r1; r++) {
        const float v = __ldg(x + r * ldx + col);
        acc += (double)v * (double)v;
    }
    partial[(size_t)blockIdx.y * d + col] = acc;
}

__global__ void __launch_bounds__(256)
sumsq_final_kernel(const double *__restrict__ partial, int M, int dsub, int n_splits, double *__restrict__ out)
{
    const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (m >= M) return;
    const int d = M * dsub;
    double s = 0.0;
    for (int p = lane; p < n_splits; p += 32)
        for (int t = 0; t < dsub; t++) s += partial[(size_t)p * d + (size_t)m * dsub + t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[m] = s;
}

}  // namespace

rb_status launch_sumsq64(const float *x, size_t n, ptrdiff_t ldx, size_t M, size_t dsub, double *out, cudaStream_t stream)
{
    if (M == 0) return RB_OK;
    const size_t d = M * dsub;
    size_t rows_per_split = 512;
    if (ceil_div(n ? n : 1, rows_per_split) > 4096) rows_per_split = ceil_div(n, (size_t)4096);
    const size_t splits = n ? ceil_div(n, rows_per_split) : 1;
    double *partial = nullptr;
    RB_CUDA_TRY(pool_malloc((void **)&partial, splits * d * sizeof(double), stream));
    sumsq_partial_kernel<<<dim3((unsigned)ceil_div(d, (size_t)256), (unsigned)splits), 256, 0, stream>>>(
        x, (long long)n, (long long)ldx, (int)d, (long long)rows_per_split, partial);
    sumsq_final_kernel<<<(unsigned)ceil_div(M, (size_t)8), 256, 0, stream>>>(partial, (int)M, (int)dsub, (int)splits, out);
    cudaFreeAsync(partial, stream);
    RB_LAUNCH_CHECK();
    return RB_OK;
}


// =========================================================================================================
// Streaming ordered update (training loops: the instances persist across iterations)
// =========================================================================================================
// The chain kernels above gather 32-byte pieces of x through L2 at ~2 TB/s whatever the prefetch depth.  A training
// loop reads the same x every iteration, so it pays to lay x out ONCE as subquantizer-major slabs
//   slabs[m][row][dsub]   (row pitch dsub floats, slab pitch n_pad * dsub floats, n_pad = n rounded up to 16)
// and to STREAM them: one CTA per (subquantizer, group of KC clusters) walks the rows in order in tiles of kStreamRows,
// double-buffered in shared memory by cp.async.bulk (contiguous copies: full HBM efficiency); per tile
//   route:  every thread takes rows of the tile and sets bit (row) of its cluster's bitmap (shared-memory atomicOr)
//           plus the word's bit in the cluster's summary word;
//   chains: lanes = (cluster, component); a lane walks ITS cluster's bits in ascending order -- the reference's row
//           order (kmeans.rs:185-189) -- and adds x[row][component] from shared memory into a register that lives
//           across tiles.  No sort, no global lists, no gathers.
// Sums are bit-identical to the other ordered paths (same sequential f32 chain per (cluster, component)).
// MEASURED (C3, one B200): 4.4 ms per iteration against 1.8 ms for sort + chains -- 20 warp-instructions per (row,
// subquantizer) pair, most of them the bit-walking loops' branches and index arithmetic (ncu: IMAD 21 %, BRA 10 %,
// ISETP 10 %, BSYNC / BSSY 12 %; FADD 2 %), half of the shared-memory wavefronts bank conflicts.  Opt-in only
// (rb_set_kmeans_update(3)); what it would need is the expansion of the bitmaps into per-cluster row lists by ONE lane
// per cluster, so that the eight component lanes of a cluster do not all walk the bits.
namespace {

constexpr int kStreamRows = 1024;    // rows per tile (32 bitmap words per cluster: one summary word)
constexpr int kStreamThreads = 512;  // 16 warps

__global__ void __launch_bounds__(256)
slab_kernel(const float *__restrict__ x, long long n, long long ldx, int M, int dsub, long long n_pad, float *__restrict__ slabs)
{
    // one block: 64 rows x all columns, staged through registers; reads are row-contiguous, writes per-slab contiguous
    const long long r0 = (long long)blockIdx.x * 64;
    const int d = M * dsub;
    for (int e = threadIdx.x; e < 64 * d; e += 256) {
        const int r = e / d, c = e - r * d;
        const long long row = r0 + r;
        if (row < n) {
            const int m = c / dsub, t = c - m * dsub;
            slabs[((long long)m * n_pad + row) * dsub + t] = __ldg(x + row * ldx + c);
        }
    }
}

template <int DSUB>
__global__ void __launch_bounds__(kStreamThreads)
ordered_stream_kernel(const float *__restrict__ slabs, long long n, long long n_pad, const uint8_t *__restrict__ codes,
                      long long code_pitch, int M, int k, int KC, float *__restrict__ packed)
{
    using namespace ptx;
    constexpr int CH = 32 / DSUB;  // clusters a warp advances at the same time
    constexpr int R = kStreamRows;
    extern __shared__ __align__(128) unsigned char smem[];
    float *sx = reinterpret_cast<float *>(smem);                               // [2][R][DSUB]
    uint8_t *sc = smem + (size_t)2 * R * DSUB * 4;                             // [2][R]
    uint32_t *bm = reinterpret_cast<uint32_t *>(sc + 2 * R);                   // [KC][32]
    uint32_t *sm = bm + (size_t)KC * 32;                                       // [KC] summary words
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + ((KC + 1) & ~1));       // [2]
    const int m = blockIdx.y, c0 = blockIdx.x * KC;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kc_cur = min(KC, k - c0);
    const float *slab = slabs + (size_t)m * n_pad * DSUB;
    const uint8_t *col = codes + (size_t)m * code_pitch;
    const long long n_tiles = (n + R - 1) / R;

    for (int i = threadIdx.x; i < KC * 32; i += kStreamThreads) bm[i] = 0u;
    for (int i = threadIdx.x; i < KC; i += kStreamThreads) sm[i] = 0u;
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    auto load_tile = [&](long long tile) {  // one thread: the tile's rows of x and their codes, contiguous in memory
        const int buf = (int)(tile & 1);
        const long long r0 = tile * R;
        const long long rows = min((long long)R, n_pad - r0);  // padded rows are readable (slab / code pitch padding)
        const uint32_t xb = (uint32_t)(rows * DSUB * 4), cb = (uint32_t)((rows + 15) / 16 * 16);
        mbar_arrive_expect_tx(&bars[buf], xb + cb);
        bulk_g2s(sx + (size_t)buf * R * DSUB, slab + r0 * DSUB, xb, &bars[buf]);
        bulk_g2s(sc + (size_t)buf * R, col + r0, cb, &bars[buf]);
    };
    if (threadIdx.x == 0 && n_tiles > 0) load_tile(0);

    // chain ownership: warp w advances the clusters w * per_warp + g * CH + cs (g = pass), lane = (cs, t)
    const int per_warp = (KC + 15) / 16;
    const int passes = (per_warp + CH - 1) / CH;
    const int cs = lane / DSUB, t = lane % DSUB;
    const bool lane_on = cs < CH;
    constexpr int kMaxPasses = 4;  // KC <= 256: per_warp <= 16, CH >= 1 ... covered by the host's choice of KC
    float acc[kMaxPasses * 4];
    unsigned cnt[kMaxPasses * 4];
#pragma unroll
    for (int g = 0; g < kMaxPasses * 4; g++) {
        acc[g] = 0.f;
        cnt[g] = 0u;
    }

    for (long long tile = 0; tile < n_tiles; tile++) {
        const int buf = (int)(tile & 1);
        // the other buffer was consumed in the previous iteration (barrier at its end): refill it now
        if (threadIdx.x == 0 && tile + 1 < n_tiles) load_tile(tile + 1);
        mbar_wait(&bars[buf], (uint32_t)((tile >> 1) & 1));
        const long long r0 = tile * R;
        const int rows = (int)min((long long)R, n - r0);
        const uint8_t *tc = sc + (size_t)buf * R;
        // ---- route ----
        for (int r = threadIdx.x; r < rows; r += kStreamThreads) {
            const int cl = (int)tc[r] - c0;
            if ((unsigned)cl < (unsigned)kc_cur) {
                atomicOr(&bm[cl * 32 + (r >> 5)], 1u << (r & 31));
                atomicOr(&sm[cl], 1u << (r >> 5));
            }
        }
        __syncthreads();
        // ---- chains ----
        const float *tx = sx + (size_t)buf * R * DSUB;
#pragma unroll
        for (int g = 0; g < kMaxPasses * 4; g++) {
            if (g >= passes) break;
            const int cl = warp * per_warp + g * CH + cs;
            const bool on = lane_on && (g * CH + cs) < per_warp && cl < kc_cur;
            uint32_t words = on ? sm[cl] : 0u;
            float a = acc[g];
            unsigned c = cnt[g];
            while (__any_sync(0xffffffffu, words != 0u)) {
                if (words != 0u) {
                    const int w = __ffs(words) - 1;
                    words &= words - 1;
                    uint32_t bits = bm[cl * 32 + w];
                    c += __popc(bits);
                    while (bits != 0u) {
                        const int b = __ffs(bits) - 1;
                        bits &= bits - 1;
                        a = __fadd_rn(a, tx[(size_t)(32 * w + b) * DSUB + t]);  // kmeans.rs:185-189: one rounded add per row
                    }
                    if (t == 0) bm[cl * 32 + w] = 0u;
                }
            }
            if (on && t == 0) sm[cl] = 0u;
            acc[g] = a;
            cnt[g] = c;
        }
        __syncthreads();  // bitmaps are clean and this tile's buffer is free
    }
    // results
    const size_t chains = (size_t)M * k * DSUB;
#pragma unroll
    for (int g = 0; g < kMaxPasses * 4; g++) {
        if (g >= passes) break;
        const int cl = warp * per_warp + g * CH + cs;
        if (lane_on && (g * CH + cs) < per_warp && cl < kc_cur) {
            const size_t mj = (size_t)m * k + c0 + cl;
            packed[mj * DSUB + t] = acc[g];
            if (t == 0) packed[chains + mj] = (float)cnt[g];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) packed[chains + (size_t)M * k + m] = 0.f;  // sum ||x||^2: callers keep it in FP64
}

int stream_cluster_groups(size_t M, size_t k)
{
    // as many (subquantizer, cluster group) CTAs as fit twice per SM, at most 8 groups, at least 16 clusters per group
    int cg = 1;
    while (cg < 8 && (size_t)(2 * cg) * M <= (size_t)2 * sm_count() && k / (size_t)(2 * cg) >= 16) cg *= 2;
    return cg;
}

}  // namespace

bool stream_update_supported(size_t k, size_t dsub)
{
    if (k > 256 || k < 16) return false;
    switch (dsub) {
    case 2: case 4: case 8: case 16: case 10: case 12: case 6: return true;  // 32 / dsub clusters per warp, tile <= 64 KB
    default: return false;
    }
}

size_t slab_floats(size_t n, size_t d) { return ((n + 15) / 16 * 16 + (size_t)kStreamRows) * d; }

rb_status launch_build_slabs(const float *x, size_t n, ptrdiff_t ldx, size_t M, size_t dsub, float *slabs, cudaStream_t stream)
{
    if (n == 0) return RB_OK;
    const size_t n_pad = (n + 15) / 16 * 16;
    RB_CUDA_TRY(cudaMemsetAsync(slabs, 0, slab_floats(n, M * dsub) * sizeof(float), stream));
    slab_kernel<<<(unsigned)ceil_div(n, (size_t)64), 256, 0, stream>>>(x, (long long)n, (long long)ldx, (int)M, (int)dsub,
                                                                        (long long)n_pad, slabs);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

rb_status launch_ordered_stream(const float *slabs, size_t n, const uint8_t *codes, size_t code_pitch, size_t M, size_t k,
                                size_t dsub, float *packed, cudaStream_t stream)
{
    const size_t len = rb_kmeans_packed_len(M, k, dsub);
    RB_CUDA_TRY(cudaMemsetAsync(packed, 0, len * sizeof(float), stream));
    if (n == 0) return RB_OK;
    const int cg = stream_cluster_groups(M, k);
    const int KC = (int)ceil_div(k, (size_t)cg);
    const size_t n_pad = (n + 15) / 16 * 16;
    const size_t smem = (size_t)2 * kStreamRows * dsub * 4 + 2 * kStreamRows + (size_t)KC * 32 * 4 + (size_t)((KC + 1) & ~1) * 4 + 16;
#define RB_STREAM(D)                                                                                                   \
    case D: {                                                                                                          \
        auto kern = ordered_stream_kernel<D>;                                                                          \
        RB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));               \
        kern<<<dim3((unsigned)cg, (unsigned)M), kStreamThreads, smem, stream>>>(slabs, (long long)n, (long long)n_pad, codes, \
                                                                                (long long)code_pitch, (int)M, (int)k, KC, packed); \
        break;                                                                                                         \
    }
    switch (dsub) {
        RB_STREAM(2) RB_STREAM(4) RB_STREAM(6) RB_STREAM(8) RB_STREAM(10) RB_STREAM(12) RB_STREAM(16)
    default:
        set_error("streaming update: subvector width %zu is not instantiated", dsub);
        return RB_ERR_UNSUPPORTED;
    }
#undef RB_STREAM
    RB_LAUNCH_CHECK();
    return RB_OK;
}

}  // namespace rb
