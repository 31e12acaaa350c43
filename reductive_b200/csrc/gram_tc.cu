// gram_tc.cu — the Gram matrices of Opq / GaussianOpq training on the tcgen05 tensor cores.
//
// Replaces, for large row counts, the FP32 CUDA-core kernel of opq.cu in
//   Covariance::covariance      src/linalg.rs:23-44   centered^T . (centered / (n - 1))
//   the Procrustes input X^T.Y^ src/pq/opq.rs:187
//   out[i, j] = (1 / b_div) * sum_r (a[r, i] - a_sub[i]) * (b[r, j] - b_sub[j]),   16 <= da, db <= 512.
//
// The contraction runs over the ROWS of two row-major matrices, i.e. both operands are "MN-major" in memory.  Two
// kernels:
//   gram_limbs_kernel  streams a matrix once: a thread reads 8 consecutive rows of one column (coalesced across the
//     lanes' columns), centres them, splits each value into two BF16 limbs (hi = bf16(v), lo = bf16(v - hi)) and
//     writes the 8 values as ONE 16-byte vector -- global layout [row group of 8][column][8 x bf16], which is exactly
//     a K-major core-matrix layout: the transposition is free, it is just where the vector is stored;
//   gram_mma_kernel    grid = (128-column tiles of a, column blocks of b (<= 160 wide), row splits), one CTA per SM.
//     A producer thread bulk-copies (cp.async.bulk) the limb slices of 16 rows per stage straight into the operand
//     layout -- no conversion, no register staging; the issuer thread runs hi.hi + hi.lo + lo.hi per stage into an
//     FP32 accumulator in TMEM (the dropped lo.lo term is below 2^-16 of |a||b|).
// FP32 accumulation in the tensor core truncates (scripts/microbench/probe.cu; measured here: a chain of 384
// accumulations leaves diagonal entries 1.2e-5 low), a bias proportional to the length of the chain.  The issuer
// therefore alternates between TWO accumulators every 256 rows (48 accumulations: < 1.5e-6); four drain warps add the
// finished one into an FP32 tile in shared memory with ordinary rounded adds while the other accumulates, so the
// tensor pipe never waits for the drain.  At the end the tile goes to a per-split partial buffer and a second kernel
// adds the splits in order (deterministic, no float atomics) and applies 1 / b_div.
#include "common.cuh"
#include "sm100_ptx.cuh"

#include <cuda_bf16.h>

namespace rb {
namespace {

using namespace ptx;

constexpr int kGmThreads = 192;     // warp 0: producer, warp 1: issuer, warps 2-5: drain (TMEM lane quarter = warp % 4)
constexpr int kGmChainSteps = 16;   // K = 16 steps (rows / 16) per accumulator turn: 256 rows
constexpr int kGmMaxBlock = 160;    // widest column block of b per CTA

// tcgen05.wait::ld that also "touches" the loaded registers, so that no use of them is scheduled before the wait
__device__ __forceinline__ void tmem_wait_ld16(uint32_t (&v)[16])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                   "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
                 :
                 : "memory");
}

// ---- limbs ---------------------------------------------------------------------------------------------------------
// out_hi / out_lo: [groups][dpad] 16-byte vectors; rows >= n and columns >= d are zero.
__global__ void __launch_bounds__(128)
gram_limbs_kernel(const float *__restrict__ x, long long n, long long ld, int d, int dpad, long long groups,
                  const float *__restrict__ sub, uint4 *__restrict__ out_hi, uint4 *__restrict__ out_lo)
{
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c >= dpad) return;
    const bool col_ok = c < d;
    const float sv = (col_ok && sub) ? __ldg(sub + c) : 0.f;
    constexpr int GP = 4;  // groups per thread and pass: 32 loads in flight
    for (long long g0 = (long long)blockIdx.y * GP; g0 < groups; g0 += (long long)gridDim.y * GP) {
        float v[GP][8];
#pragma unroll
        for (int q = 0; q < GP; q++) {
            const long long r = (g0 + q) * 8;
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const bool ok = col_ok && r + e < n;
                const float xv = ok ? __ldg(x + (r + e) * ld + c) : 0.f;
                v[q][e] = ok ? __fsub_rn(xv, sv) : 0.f;  // `centered` (linalg.rs:30-31); padding contributes zero
            }
        }
#pragma unroll
        for (int q = 0; q < GP; q++) {
            if (g0 + q >= groups) break;
            uint32_t h[4], l[4];
#pragma unroll
            for (int pr = 0; pr < 4; pr++) {
                const __nv_bfloat162 hh = __floats2bfloat162_rn(v[q][2 * pr], v[q][2 * pr + 1]);
                const float2 hf = __bfloat1622float2(hh);
                const __nv_bfloat162 ll = __floats2bfloat162_rn(__fsub_rn(v[q][2 * pr], hf.x), __fsub_rn(v[q][2 * pr + 1], hf.y));
                h[pr] = *reinterpret_cast<const uint32_t *>(&hh);
                l[pr] = *reinterpret_cast<const uint32_t *>(&ll);
            }
            const size_t o = (size_t)(g0 + q) * dpad + c;
            out_hi[o] = make_uint4(h[0], h[1], h[2], h[3]);
            out_lo[o] = make_uint4(l[0], l[1], l[2], l[3]);
        }
    }
}

// ---- MMA -----------------------------------------------------------------------------------------------------------
struct GramMmaParams {
    const uint4 *a_hi, *a_lo, *b_hi, *b_lo;  // limb arrays, [groups][dpad_a] / [groups][dpad_b]
    int dpad_a, dpad_b;
    long long steps;        // 16-row steps in total (limb arrays hold 2 * steps groups)
    long long steps_per_split;
    int da, db, nw, stages;
    float *partial;         // [splits][db][da]
};

__global__ void __launch_bounds__(kGmThreads, 1) gram_mma_kernel(const __grid_constant__ GramMmaParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nw = p.nw;
    const int a_limb = 2 * 128 * 16, b_limb = 2 * nw * 16;  // bytes of one limb of one stage (two row groups)
    const int stage_bytes = 2 * a_limb + 2 * b_limb;
    float *tile = reinterpret_cast<float *>(smem);          // [nw][128]
    unsigned char *stages = smem + (size_t)nw * 128 * sizeof(float);
    uint64_t *bars = reinterpret_cast<uint64_t *>(stages + (size_t)p.stages * stage_bytes);
    uint64_t *full = bars, *empty = bars + 8, *acc_full = bars + 16, *acc_empty = bars + 18;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 20);

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 4);
        }
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

    const int i0 = blockIdx.x * 128, j0 = blockIdx.y * nw;
    const long long t0 = (long long)blockIdx.z * p.steps_per_split;
    const long long t1 = min(p.steps, t0 + p.steps_per_split);
    const long long n_steps = t1 > t0 ? t1 - t0 : 0;
    const long long n_turns = (n_steps + kGmChainSteps - 1) / kGmChainSteps;

    if (warp == 0) {
        // ===================== producer: bulk copies of the limb slices into the operand layout =====================
        // lanes 0-7 issue one copy each (a single thread issuing all eight was the bottleneck: ~1 100 clk per stage);
        // lane 0 also posts the byte count -- a copy that completes before the count is posted only drives the
        // transaction count negative for a moment, the phase cannot complete before lane 0's arrival
        const int q = (lane >> 2) & 1, kind = lane & 3;  // row group of the stage, (A hi, A lo, B hi, B lo)
        const uint4 *src = kind == 0 ? p.a_hi : kind == 1 ? p.a_lo : kind == 2 ? p.b_hi : p.b_lo;
        const size_t pitch = kind < 2 ? (size_t)p.dpad_a : (size_t)p.dpad_b;
        const size_t col = kind < 2 ? (size_t)i0 : (size_t)j0;
        const uint32_t bytes = kind < 2 ? 2048u : (uint32_t)nw * 16u;
        const int dst_off = (kind == 0 ? 0 : kind == 1 ? a_limb : kind == 2 ? 2 * a_limb : 2 * a_limb + b_limb) + q * (int)bytes;
        for (long long t = 0; t < n_steps; t++) {
            const int s = (int)(t % p.stages);
            mbar_wait(&empty[s], (uint32_t)(((t / p.stages) & 1) ^ 1));
            unsigned char *st = stages + (size_t)s * stage_bytes;
            if (lane == 0) mbar_arrive_expect_tx(&full[s], (uint32_t)stage_bytes);
            if (lane < 8) bulk_g2s(st + dst_off, src + (size_t)(2 * (t0 + t) + q) * pitch + col, bytes, &full[s]);
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = idesc_f16(128, (uint32_t)nw, 1);
        for (long long t = 0; t < n_steps; t++) {
            const int s = (int)(t % p.stages);
            const long long turn = t / kGmChainSteps;
            const int ab = (int)(turn & 1);
            const bool turn_start = t % kGmChainSteps == 0;
            if (turn_start) mbar_wait(&acc_empty[ab], (uint32_t)(((turn >> 1) & 1) ^ 1));  // its last drain has read it
            mbar_wait(&full[s], (uint32_t)((t / p.stages) & 1));
            tc_fence_after();
            if (lane == 0) {
                const uint32_t sa = smem_u32(stages + (size_t)s * stage_bytes);
                const uint64_t a_hi = smem_desc_kmajor(sa, 2048, 128), a_lo = smem_desc_kmajor(sa + a_limb, 2048, 128);
                const uint64_t b_hi = smem_desc_kmajor(sa + 2 * a_limb, (uint32_t)nw * 16, 128);
                const uint64_t b_lo = smem_desc_kmajor(sa + 2 * a_limb + b_limb, (uint32_t)nw * 16, 128);
                const uint32_t d = tmem_base + (uint32_t)ab * 256u;
                mma_f16_ss(d, a_hi, b_hi, idesc, turn_start ? 0u : 1u);
                mma_f16_ss(d, a_hi, b_lo, idesc, 1u);
                mma_f16_ss(d, a_lo, b_hi, idesc, 1u);
                tc_commit(&empty[s]);
                if ((t + 1) % kGmChainSteps == 0 || t + 1 == n_steps) tc_commit(&acc_full[ab]);
            }
            __syncwarp();
        }
    } else {
        // ===================== drain: TMEM lane = column i0 + row of a; tile[j][row] += accumulator =====================
        const int q4 = warp & 3, row = q4 * 32 + lane;
        for (long long turn = 0; turn < n_turns; turn++) {
            const int ab = (int)(turn & 1);
            __syncwarp();
            mbar_wait(&acc_full[ab], (uint32_t)((turn >> 1) & 1));
            __syncwarp();
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)ab * 256u;
            for (int c0 = 0; c0 < nw; c0 += 16) {  // nw is a multiple of 16
                uint32_t w[16];
                tmem_ld16(taddr + c0, w);
                tmem_wait_ld16(w);
                if (turn == 0) {
#pragma unroll
                    for (int k = 0; k < 16; k++) tile[(c0 + k) * 128 + row] = __uint_as_float(w[k]);
                } else {
#pragma unroll
                    for (int k = 0; k < 16; k++)
                        tile[(c0 + k) * 128 + row] = __fadd_rn(tile[(c0 + k) * 128 + row], __uint_as_float(w[k]));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[ab]);
        }
        // this split's tile: partial[split][j][i], coalesced over the lanes' rows
        const int i = i0 + row;
        if (i < p.da) {
            float *dst = p.partial + (size_t)blockIdx.z * p.db * p.da + i;
            for (int j = 0; j < nw; j++)
                if (j0 + j < p.db) dst[(size_t)(j0 + j) * p.da] = n_turns ? tile[j * 128 + row] : 0.f;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

__global__ void gram_tc_final_kernel(const float *__restrict__ partial, int da, int db, int n_splits, float b_div, float *__restrict__ out)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)da * db) return;
    const int i = (int)(e / db), j = (int)(e % db);
    float s = 0.f;
    for (int q = 0; q < n_splits; q++) s += partial[((size_t)q * db + j) * da + i];  // fixed order
    out[e] = b_div != 1.f ? __fdiv_rn(s, b_div) : s;
}

}  // namespace

bool gram_tensor_supported(const float *a, ptrdiff_t lda, const float *b, ptrdiff_t ldb, size_t n, size_t da, size_t db)
{
    (void)a;
    (void)b;
    return n >= 4096 && da >= 16 && db >= 16 && da <= 512 && db <= 512 && lda >= (ptrdiff_t)da && ldb >= (ptrdiff_t)db;
}

rb_status launch_gram_tensor(const float *a, ptrdiff_t lda, const float *b, ptrdiff_t ldb, size_t n, size_t da, size_t db,
                             const float *a_sub, const float *b_sub, float b_div, float *out, cudaStream_t stream)
{
    const size_t m_tiles = ceil_div(da, (size_t)128);
    const size_t npad = ceil_div(db, (size_t)16) * 16;
    const size_t n_blocks = ceil_div(npad, (size_t)kGmMaxBlock);
    const size_t nw = ceil_div(ceil_div(npad, n_blocks), (size_t)16) * 16;
    const size_t dpad_a = m_tiles * 128, dpad_b = n_blocks * nw;
    const bool same = a == b && lda == ldb && da == db && a_sub == b_sub;
    const size_t dpad_same = dpad_a > dpad_b ? dpad_a : dpad_b;
    const int dpa = (int)(same ? dpad_same : dpad_a);

    // The rows go through in passes of at most pass_rows (the limb arrays of a pass: 32 bytes per element, i.e. 8 GB
    // for 4 M rows x 512 columns and operand); every pass fills its own slots of the partial buffer and the final
    // kernel adds all of them in order.  RB_GRAM_PASS_ROWS: development aid (tests force several passes).
    const char *pass_env = getenv("RB_GRAM_PASS_ROWS");
    const long long pass_req = pass_env ? atoll(pass_env) : 0;
    const size_t pass_rows = pass_req >= 4096 ? (size_t)pass_req / 16 * 16 : (size_t)4 << 20;
    const size_t n_pass = ceil_div(n, pass_rows);
    const size_t steps_full = ceil_div(n < pass_rows ? n : pass_rows, (size_t)16);
    size_t splits_max = (size_t)sm_count() / (m_tiles * n_blocks);
    if (splits_max < 1) splits_max = 1;

    GramMmaParams p;
    p.da = (int)da;
    p.db = (int)db;
    p.nw = (int)nw;
    const size_t stage_bytes = 2 * (2 * 128 * 16) + 2 * (2 * nw * 16);
    const size_t tile_bytes = nw * 128 * sizeof(float);
    p.stages = (int)std::min<size_t>(8, (226 * 1024 - tile_bytes - 256) / stage_bytes);
    if (p.stages < 2) {
        set_error("gram_tc: shared-memory plan does not fit (nw = %zu)", nw);
        return RB_ERR_UNSUPPORTED;
    }
    const size_t smem = tile_bytes + (size_t)p.stages * stage_bytes + 256;

    uint4 *limbs = nullptr;
    float *partial = nullptr;
    const size_t groups_full = 2 * steps_full;
    const size_t a_vecs = groups_full * (size_t)dpa, b_vecs = same ? 0 : groups_full * dpad_b;
    RB_CUDA_TRY(pool_malloc((void **)&limbs, 2 * (a_vecs + b_vecs) * sizeof(uint4), stream));
    rb_status st = [&]() -> rb_status {
        RB_CUDA_TRY(pool_malloc((void **)&partial, n_pass * splits_max * da * db * sizeof(float), stream));
        RB_CUDA_TRY(cudaFuncSetAttribute(gram_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        uint4 *a_hi = limbs, *a_lo = limbs + a_vecs, *b_hi = limbs + 2 * a_vecs, *b_lo = limbs + 2 * a_vecs + b_vecs;
        size_t slots = 0;  // partial tiles written so far
        for (size_t ps = 0; ps < n_pass; ps++) {
            const size_t r0 = ps * pass_rows, rows = n - r0 < pass_rows ? n - r0 : pass_rows;
            const size_t steps = ceil_div(rows, (size_t)16), groups = 2 * steps;
            const unsigned gy = (unsigned)std::min<size_t>(ceil_div(groups, (size_t)4), (size_t)sm_count() * 16);
            gram_limbs_kernel<<<dim3((unsigned)ceil_div((size_t)dpa, (size_t)128), gy), 128, 0, stream>>>(
                a + (ptrdiff_t)r0 * lda, (long long)rows, (long long)lda, (int)da, dpa, (long long)groups, a_sub, a_hi, a_lo);
            RB_LAUNCH_CHECK();
            if (same) {
                p.b_hi = a_hi;
                p.b_lo = a_lo;
                p.dpad_a = p.dpad_b = dpa;
            } else {
                gram_limbs_kernel<<<dim3((unsigned)ceil_div(dpad_b, (size_t)128), gy), 128, 0, stream>>>(
                    b + (ptrdiff_t)r0 * ldb, (long long)rows, (long long)ldb, (int)db, (int)dpad_b, (long long)groups, b_sub, b_hi,
                    b_lo);
                RB_LAUNCH_CHECK();
                p.b_hi = b_hi;
                p.b_lo = b_lo;
                p.dpad_a = (int)dpad_a;
                p.dpad_b = (int)dpad_b;
            }
            p.a_hi = a_hi;
            p.a_lo = a_lo;
            p.steps = (long long)steps;
            size_t splits = splits_max;
            if (splits > ceil_div(steps, (size_t)64)) splits = ceil_div(steps, (size_t)64);
            p.steps_per_split = (long long)ceil_div(steps, splits);
            splits = ceil_div(steps, (size_t)p.steps_per_split);
            p.partial = partial + slots * da * db;
            gram_mma_kernel<<<dim3((unsigned)m_tiles, (unsigned)n_blocks, (unsigned)splits), kGmThreads, smem, stream>>>(p);
            RB_LAUNCH_CHECK();
            slots += splits;
        }
        gram_tc_final_kernel<<<(unsigned)ceil_div(da * db, (size_t)256), 256, 0, stream>>>(partial, (int)da, (int)db, (int)slots,
                                                                                           b_div, out);
        RB_LAUNCH_CHECK();
        return RB_OK;
    }();
    if (partial) cudaFreeAsync(partial, stream);
    cudaFreeAsync(limbs, stream);
    return st;
}

}  // namespace rb
