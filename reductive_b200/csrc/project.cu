// project.cu — the OPQ / GaussianOpq rotation as an FP32 GEMM in the reference's accumulation order (sm_100a).
//
// Replaces  x.dot(projection)                 src/pq/pq.rs:276   (encode:  rx = X . R)
//           reconstructions.dot(&projection.t())  src/pq/pq.rs:324   (decode:  X^ = Y^ . R^T)
// Both go through matrixmultiply::sgemm in the reference: every output element is ONE accumulator updated
// by a fused multiply-add sequentially over k, with K cut in blocks of kc = 256 whose partial results are
// combined by a plain add.  This kernel keeps exactly that order per output element (a thread never splits
// K), so rx is bit-identical to the oracle's and the codes that follow are bit-exact end to end.
//
// Bound: FP32 FMA pipe (2*d^2 flop per row).  project_fast_kernel is a register-tiled SIMT SGEMM (128 x 64 block
// tile, 8 x 4 per thread, K steps of 16 through a 3-stage cp.async ring, 16-byte global and shared accesses);
// project_kernel is the general-stride fallback.  Large aligned batches go through the tcgen05 two-limb GEMM in
// project_tc.cu instead (not bit-identical: decode uses it within 1e-5, encode re-rotates what it cannot certify
// with the order implemented here); this file remains the exact path for small batches, OPQ training and
// rb_project_rows.
#include "common.cuh"

namespace rb {

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

template <bool TR>
__global__ void __launch_bounds__(256)
project_kernel(const float *__restrict__ x, long long n, int d, long long rsx, long long csx,
               const float *__restrict__ r, float *__restrict__ out, long long ldo)
{
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long row0 = (long long)blockIdx.x * BM;
    const int col0 = blockIdx.y * BN;

    float acc[4][4], total[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = total[i][j] = 0.f;
    bool have = false;

    for (int t0 = 0; t0 < d; t0 += BK) {
#pragma unroll
        for (int l = 0; l < (BM * BK) / 256; l++) {
            const int idx = tid + l * 256;
            const int ak = idx % BK, ar = idx / BK;
            const long long row = row0 + ar;
            const int t = t0 + ak;
            As[ak][ar] = (row < n && t < d) ? __ldg(x + row * rsx + (long long)t * csx) : 0.f;
        }
#pragma unroll
        for (int l = 0; l < (BN * BK) / 256; l++) {
            const int idx = tid + l * 256;
            int bk, bc;
            if (TR) { bk = idx % BK; bc = idx / BK; } else { bc = idx % BN; bk = idx / BN; }
            const int t = t0 + bk, col = col0 + bc;
            float v = 0.f;
            if (t < d && col < d) v = TR ? __ldg(r + (size_t)col * d + t) : __ldg(r + (size_t)t * d + col);
            Bs[bk][bc] = v;
        }
        __syncthreads();
        const int kmax = min(BK, d - t0);
        for (int kk = 0; kk < kmax; kk++) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = __fmaf_rn(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
        // matrixmultiply kc = 256: first block C = AB, later blocks C = C + AB
        if (((t0 + BK) & 255) == 0 || t0 + BK >= d) {
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    total[i][j] = have ? __fadd_rn(total[i][j], acc[i][j]) : acc[i][j];
                    acc[i][j] = 0.f;
                }
            have = true;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const long long row = row0 + ty * 4 + i;
        if (row >= n) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int col = col0 + tx * 4 + j;
            if (col < d) out[row * ldo + col] = total[i][j];
        }
    }
}

// ---- fast path: unit column stride, 16-byte aligned rows, d % 4 == 0, R not transposed ------------------------
// 64 x 64 block tile (PROJ_PM), 16 x 16 threads, thread tile 4 rows (ty + 16 i) x 4 columns; K steps of 16 streamed through a
// 3-stage shared-memory ring by cp.async (16 bytes each), operands read back as 128-bit vectors: A[row][k..k+3] (row
// pitch 20 floats: the two rows a warp touches fall in different banks) and B[k][col..col+3].
#ifndef PROJ_PM
#define PROJ_PM 64
#endif
constexpr int PM = PROJ_PM, PN = 64, PK = 16, PSTAGES = 3, PA_PITCH = PK + 4, PB_PITCH = PN + 4;
constexpr int PR = PM / 16;  // rows per thread

__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src, bool valid)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    const int bytes = valid ? 16 : 0;  // src-size 0: the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(256)
project_fast_kernel(const float *__restrict__ x, long long n, int d, long long ldx, const float *__restrict__ r,
                    float *__restrict__ out, long long ldo)
{
    __shared__ __align__(16) float As[PSTAGES][PM][PA_PITCH];
    __shared__ __align__(16) float Bs[PSTAGES][PK][PB_PITCH];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long row0 = (long long)blockIdx.x * PM;
    const int col0 = blockIdx.y * PN;
    const int n_steps = (d + PK - 1) / PK;

    auto issue = [&](int step) {
        if (step < n_steps) {
            const int t0 = step * PK, st = step % PSTAGES;
#pragma unroll
            for (int l = 0; l < PM / 64; l++) {
                const int idx = tid + l * 256;
                const int ar = idx >> 2, c4 = idx & 3;
                const long long row = row0 + ar;
                const int t = t0 + 4 * c4;
                const bool ok = row < n && t < d;
                cp_async16(&As[st][ar][4 * c4], ok ? x + row * ldx + t : x, ok);
            }
            const int bk = tid >> 4, c4 = tid & 15;
            const int t = t0 + bk, col = col0 + 4 * c4;
            const bool ok = t < d && col < d;
            cp_async16(&Bs[st][bk][4 * c4], ok ? r + (size_t)t * d + col : r, ok);
        }
        cp_async_commit();  // one group per step, empty past the end, so the wait counts stay uniform
    };

    float acc[PR][4], total[PR][4];
#pragma unroll
    for (int i = 0; i < PR; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = total[i][j] = 0.f;
    bool have = false;

    issue(0);
    issue(1);
    for (int step = 0; step < n_steps; step++) {
        issue(step + 2);
        cp_async_wait<2>();  // this step's group has landed (two younger groups may be in flight)
        __syncthreads();
        const int st = step % PSTAGES, t0 = step * PK;
        auto quad = [&](int kq, int kn) {  // kn (<= 4) consecutive k starting at 4 * kq, in order
            float4 a[PR];
#pragma unroll
            for (int i = 0; i < PR; i++) a[i] = *reinterpret_cast<const float4 *>(&As[st][ty + 16 * i][4 * kq]);
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
                if (kk < kn) {
                    const float4 b = *reinterpret_cast<const float4 *>(&Bs[st][4 * kq + kk][tx * 4]);
                    const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int i = 0; i < PR; i++) {
                        const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
#pragma unroll
                        for (int j = 0; j < 4; j++) acc[i][j] = __fmaf_rn(av, bb[j], acc[i][j]);
                    }
                }
            }
        };
        if (t0 + PK <= d) {
#pragma unroll
            for (int kq = 0; kq < PK / 4; kq++) quad(kq, 4);
        } else {  // ragged last step: exactly the reference's d - t0 updates (a padded fma could turn -0 into +0)
            const int left = d - t0;
            for (int kq = 0; 4 * kq < left; kq++) quad(kq, min(4, left - 4 * kq));
        }
        // matrixmultiply kc = 256: first block C = AB, later blocks C = C + AB
        if (((t0 + PK) & 255) == 0 || step + 1 == n_steps) {
#pragma unroll
            for (int i = 0; i < PR; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    total[i][j] = have ? __fadd_rn(total[i][j], acc[i][j]) : acc[i][j];
                    acc[i][j] = 0.f;
                }
            have = true;
        }
        __syncthreads();  // everyone is done with this stage before it is refilled two steps later
    }
    const int col = col0 + tx * 4;
    if (col < d) {
#pragma unroll
        for (int i = 0; i < PR; i++) {
            const long long row = row0 + ty + 16 * i;
            if (row < n)
                *reinterpret_cast<float4 *>(out + row * ldo + col) = make_float4(total[i][0], total[i][1], total[i][2], total[i][3]);
        }
    }
}

__global__ void pack_rows_kernel(const float *__restrict__ src, long long n, long long d, long long rs, long long cs,
                                 float *__restrict__ dst)
{
    const long long total = n * d;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / d, c = i - r * d;
        dst[i] = src[r * rs + c * cs];
    }
}

__global__ void unpack_rows_kernel(const float *__restrict__ src, long long n, long long d, float *__restrict__ dst,
                                   long long rs, long long cs)
{
    const long long total = n * d;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / d, c = i - r * d;
        dst[r * rs + c * cs] = src[i];
    }
}

}  // namespace

rb_status launch_project(const float *x, size_t n, size_t d, ptrdiff_t rsx, ptrdiff_t csx, const float *r,
                         int transpose_r, float *out, cudaStream_t stream)
{
    if (n == 0 || d == 0) return RB_OK;
    const bool aligned = csx == 1 && d % 4 == 0 && rsx % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                         (reinterpret_cast<uintptr_t>(r) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    if (aligned && !transpose_r) {  // decode passes a pre-transposed copy of R instead of transpose_r (cabi.cu)
        dim3 fgrid((unsigned)ceil_div(n, PM), (unsigned)ceil_div(d, PN));
        project_fast_kernel<<<fgrid, 256, 0, stream>>>(x, (long long)n, (int)d, (long long)rsx, r, out, (long long)d);
        RB_LAUNCH_CHECK();
        return RB_OK;
    }
    dim3 grid((unsigned)ceil_div(n, BM), (unsigned)ceil_div(d, BN));
    if (transpose_r)
        project_kernel<true><<<grid, 256, 0, stream>>>(x, (long long)n, (int)d, (long long)rsx, (long long)csx, r, out,
                                                       (long long)d);
    else
        project_kernel<false><<<grid, 256, 0, stream>>>(x, (long long)n, (int)d, (long long)rsx, (long long)csx, r,
                                                        out, (long long)d);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

rb_status launch_pack_rows(const float *src, size_t n, size_t d, ptrdiff_t rs, ptrdiff_t cs, float *dst,
                           cudaStream_t stream)
{
    if (n == 0 || d == 0) return RB_OK;
    pack_rows_kernel<<<(unsigned)sm_count() * 8, 256, 0, stream>>>(src, (long long)n, (long long)d, (long long)rs, (long long)cs, dst);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

rb_status launch_unpack_rows(const float *src, size_t n, size_t d, float *dst, ptrdiff_t rs, ptrdiff_t cs,
                             cudaStream_t stream)
{
    if (n == 0 || d == 0) return RB_OK;
    unpack_rows_kernel<<<(unsigned)sm_count() * 8, 256, 0, stream>>>(src, (long long)n, (long long)d, dst, (long long)rs,
                                                    (long long)cs);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

}  // namespace rb
