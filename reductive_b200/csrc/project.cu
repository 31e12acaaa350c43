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
// tile, 8 x 4 per thread, K steps of 16 double-buffered through shared memory, 16-byte global and shared accesses);
// project_kernel is the general-stride fallback.  A tcgen05 3xBF16 split would be several times faster still but
// not bit-identical; it is listed as follow-up work in DESIGN.md.
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

// ---- fast path: unit column stride, 16-byte aligned rows, d % 4 == 0 ----------------------------------------
constexpr int PM = 128, PN = 64, PK = 16;
#ifndef PROJ_BLOCKS
#define PROJ_BLOCKS 1
#endif

template <bool TR>
__global__ void __launch_bounds__(256, PROJ_BLOCKS)
project_fast_kernel(const float *__restrict__ x, long long n, int d, long long ldx, const float *__restrict__ r,
                    float *__restrict__ out, long long ldo)
{
    __shared__ __align__(16) float As[2][PK][PM + 4];  // [k][row]
    __shared__ __align__(16) float Bs[2][PK][PN + 4];  // [k][col]
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long row0 = (long long)blockIdx.x * PM;
    const int col0 = blockIdx.y * PN;

    // global -> register staging of one K step (A: two float4 per thread, B: one)
    float4 pa[2], pb;
    auto fetch = [&](int t0) {
#pragma unroll
        for (int l = 0; l < 2; l++) {
            const int idx = tid + l * 256;
            const int ar = idx >> 2, c4 = idx & 3;
            const long long row = row0 + ar;
            const int t = t0 + 4 * c4;
            pa[l] = (row < n && t < d) ? __ldg(reinterpret_cast<const float4 *>(x + row * ldx + t)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (TR) {  // B[k][c] = R[col0 + c][t0 + k]: 16 bytes along k
            const int bc = tid >> 2, k4 = tid & 3;
            const int col = col0 + bc, t = t0 + 4 * k4;
            pb = (col < d && t < d) ? __ldg(reinterpret_cast<const float4 *>(r + (size_t)col * d + t)) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {   // B[k][c] = R[t0 + k][col0 + c]: 16 bytes along c
            const int bk = tid >> 4, c4 = tid & 15;
            const int t = t0 + bk, col = col0 + 4 * c4;
            pb = (t < d && col < d) ? __ldg(reinterpret_cast<const float4 *>(r + (size_t)t * d + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int l = 0; l < 2; l++) {
            const int idx = tid + l * 256;
            const int ar = idx >> 2, c4 = idx & 3;
            As[buf][4 * c4 + 0][ar] = pa[l].x;
            As[buf][4 * c4 + 1][ar] = pa[l].y;
            As[buf][4 * c4 + 2][ar] = pa[l].z;
            As[buf][4 * c4 + 3][ar] = pa[l].w;
        }
        if (TR) {
            const int bc = tid >> 2, k4 = tid & 3;
            Bs[buf][4 * k4 + 0][bc] = pb.x;
            Bs[buf][4 * k4 + 1][bc] = pb.y;
            Bs[buf][4 * k4 + 2][bc] = pb.z;
            Bs[buf][4 * k4 + 3][bc] = pb.w;
        } else {
            const int bk = tid >> 4, c4 = tid & 15;
            *reinterpret_cast<float4 *>(&Bs[buf][bk][4 * c4]) = pb;
        }
    };

    float acc[8][4], total[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = total[i][j] = 0.f;
    bool have = false;

    fetch(0);
    stash(0);
    __syncthreads();
    int buf = 0;
    for (int t0 = 0; t0 < d; t0 += PK, buf ^= 1) {
        const bool more = t0 + PK < d;
        if (more) fetch(t0 + PK);
        auto step = [&](int kk) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][kk][ty * 8 + 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&Bs[buf][kk][tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = __fmaf_rn(a[i], bb[j], acc[i][j]);
        };
        if (t0 + PK <= d) {
#pragma unroll
            for (int kk = 0; kk < PK; kk++) step(kk);
        } else {  // ragged last step: exactly the reference's d - t0 updates (a padded fma could turn -0 into +0)
            for (int kk = 0; kk < d - t0; kk++) step(kk);
        }
        // matrixmultiply kc = 256: first block C = AB, later blocks C = C + AB
        if (((t0 + PK) & 255) == 0 || !more) {
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    total[i][j] = have ? __fadd_rn(total[i][j], acc[i][j]) : acc[i][j];
                    acc[i][j] = 0.f;
                }
            have = true;
        }
        if (more) stash(buf ^ 1);
        __syncthreads();
    }
    const int col = col0 + tx * 4;
    if (col < d) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const long long row = row0 + ty * 8 + i;
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
    if (aligned) {
        dim3 fgrid((unsigned)ceil_div(n, PM), (unsigned)ceil_div(d, PN));
        if (transpose_r)
            project_fast_kernel<true><<<fgrid, 256, 0, stream>>>(x, (long long)n, (int)d, (long long)rsx, r, out, (long long)d);
        else
            project_fast_kernel<false><<<fgrid, 256, 0, stream>>>(x, (long long)n, (int)d, (long long)rsx, r, out, (long long)d);
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
    pack_rows_kernel<<<148 * 8, 256, 0, stream>>>(src, (long long)n, (long long)d, (long long)rs, (long long)cs, dst);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

rb_status launch_unpack_rows(const float *src, size_t n, size_t d, float *dst, ptrdiff_t rs, ptrdiff_t cs,
                             cudaStream_t stream)
{
    if (n == 0 || d == 0) return RB_OK;
    unpack_rows_kernel<<<148 * 8, 256, 0, stream>>>(src, (long long)n, (long long)d, dst, (long long)rs,
                                                    (long long)cs);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

}  // namespace rb
