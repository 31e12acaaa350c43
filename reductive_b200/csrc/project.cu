// project.cu — the OPQ / GaussianOpq rotation as an FP32 GEMM in the reference's accumulation order (sm_100a).
//
// Replaces  x.dot(projection)                 src/pq/pq.rs:276   (encode:  rx = X . R)
//           reconstructions.dot(&projection.t())  src/pq/pq.rs:324   (decode:  X^ = Y^ . R^T)
// Both go through matrixmultiply::sgemm in the reference: every output element is ONE accumulator updated
// by a fused multiply-add sequentially over k, with K cut in blocks of kc = 256 whose partial results are
// combined by a plain add.  This kernel keeps exactly that order per output element (a thread never splits
// K), so rx is bit-identical to the oracle's and the codes that follow are bit-exact end to end.
//
// Bound: FP32 FMA pipe (2*d^2 flop per row).  A tcgen05 3xBF16 split would be ~10x faster but not
// bit-identical; it is listed as follow-up work in DESIGN.md.
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
