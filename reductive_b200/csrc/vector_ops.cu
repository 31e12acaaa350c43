// vector_ops.cu — single-vector quantize / reconstruct (latency paths, one block each; sm_100a).
//
// Replaces  QuantizeVector::quantize_vector  src/pq/pq.rs:285-298 -> primitives::quantize  primitives.rs:14-49
//           -> kmeans::cluster_assignment  kmeans.rs:111-126 -> SquaredEuclideanDistance<Ix2> for Ix1  linalg.rs:118-148
//           Reconstruct::reconstruct_into    src/pq/pq.rs:329-343 -> primitives::reconstruct_into primitives.rs:110-148
// The arithmetic differs from the batch path on purpose (as in the reference): the mat-vec goes through
// ndarray's row-wise dot (unrolled_dot for contiguous operands, plain sequential sum otherwise), not the
// GEMM micro-kernel, so near-ties may legitimately resolve differently from quantize_batch.
#include "common.cuh"

namespace rb {

namespace {

constexpr int kVecThreads = 256;

__global__ void __launch_bounds__(kVecThreads)
quantize_vector_kernel(const float *__restrict__ quantizers, const float *__restrict__ cs_all, int M, int k, int dsub,
                       const float *__restrict__ projection, const float *__restrict__ x, long long sx, void *codes,
                       int code_width, long long cstride, float *__restrict__ v)
{
    const int d = M * dsub;
    // pq.rs:293: x.dot(projection) == projection.t().dot(x): one row.dot(x) per (non-contiguous) column of R
    for (int j = threadIdx.x; j < d; j += kVecThreads) {
        if (projection) {
            float sum = 0.f;
            for (int i = 0; i < d; i++) sum = __fadd_rn(sum, __fmul_rn(projection[(size_t)i * d + j], x[(long long)i * sx]));
            v[j] = sum;
        } else {
            v[j] = x[(long long)j * sx];
        }
    }
    __syncthreads();
    // a strided input view is not a slice, so ndarray takes the sequential dot for it (see oracle.c)
    const bool seq = (!projection && sx != 1 && d > 1 && dsub > 1);

    __shared__ float s_val[kVecThreads / 32];
    __shared__ int s_idx[kVecThreads / 32];

    for (int m = 0; m < M; m++) {
        const float *sub = v + (size_t)m * dsub;
        const float *qm = quantizers + (size_t)m * k * dsub;
        float xs;
        if (seq) {
            xs = 0.f;
            for (int t = 0; t < dsub; t++) xs = __fadd_rn(xs, __fmul_rn(sub[t], sub[t]));
        } else {
            xs = unrolled_dot_dev(dsub, [&](int i) { return sub[i]; }, [&](int i) { return sub[i]; });  // linalg.rs:136
        }
        int best = -1;
        float bv = 0.f;
        for (int j = threadIdx.x; j < k; j += kVecThreads) {
            const float *cj = qm + (size_t)j * dsub;
            float dp;
            if (seq) {
                dp = 0.f;
                for (int t = 0; t < dsub; t++) dp = __fadd_rn(dp, __fmul_rn(cj[t], sub[t]));
            } else {
                dp = unrolled_dot_dev(dsub, [&](int i) { return cj[i]; }, [&](int i) { return sub[i]; });  // linalg.rs:141
            }
            const float dist = ref_distance(xs, cs_all[(size_t)m * k + j], dp);  // linalg.rs:143
            if (best < 0 || of_less(dist, bv)) {
                bv = dist;
                best = j;
            }
        }
        // block argmin: smaller distance, then smaller index (min_by_key keeps the first minimum)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
            const int oi = __shfl_xor_sync(0xffffffffu, best, off);
            if (oi >= 0 && (best < 0 || of_less(ov, bv) || (!of_less(bv, ov) && oi < best))) {
                bv = ov;
                best = oi;
            }
        }
        if ((threadIdx.x & 31) == 0) {
            s_val[threadIdx.x >> 5] = bv;
            s_idx[threadIdx.x >> 5] = best;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < kVecThreads / 32; w++) {
                const float ov = s_val[w];
                const int oi = s_idx[w];
                if (oi >= 0 && (best < 0 || of_less(ov, bv) || (!of_less(bv, ov) && oi < best))) {
                    bv = ov;
                    best = oi;
                }
            }
            store_code(codes, code_width, (long long)m * cstride, (unsigned)best);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kVecThreads)
reconstruct_vector_kernel(const float *__restrict__ quantizers, int M, int k, int dsub,
                          const float *__restrict__ projection, const void *__restrict__ codes, int code_width,
                          long long cstride, float *__restrict__ out, long long ostride, float *__restrict__ tmp,
                          int *__restrict__ err_flag)
{
    const int d = M * dsub;
    __shared__ int bad;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < d; i += kVecThreads) {
        const int m = i / dsub, t = i - m * dsub;
        const unsigned long long code = load_code(codes, code_width, (long long)m * cstride);
        if (code >= (unsigned long long)k) {
            bad = 1;
            tmp[i] = 0.f;
        } else {
            tmp[i] = quantizers[((size_t)m * k + code) * dsub + t];  // primitives.rs:146
        }
    }
    __syncthreads();
    if (bad) {
        if (threadIdx.x == 0) atomicExch(err_flag, 1);
        return;
    }
    for (int i = threadIdx.x; i < d; i += kVecThreads) {
        float r = tmp[i];
        if (projection) {
            // pq.rs:340: reconstruction.dot(&projection.t()) == projection.dot(reconstruction): contiguous rows
            const float *row = projection + (size_t)i * d;
            r = unrolled_dot_dev(d, [&](int t) { return row[t]; }, [&](int t) { return tmp[t]; });
        }
        out[(long long)i * ostride] = r;
    }
}

}  // namespace

rb_status launch_quantize_vector(const DeviceCodebook &cb, const float *projection, const float *x, ptrdiff_t sx,
                                 void *codes, int code_width, ptrdiff_t cstride, float *scratch, cudaStream_t stream)
{
    quantize_vector_kernel<<<1, kVecThreads, 0, stream>>>(cb.quantizers, cb.cs, (int)cb.M, (int)cb.k, (int)cb.dsub,
                                                          projection, x, (long long)sx, codes, code_width,
                                                          (long long)cstride, scratch);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

rb_status launch_reconstruct_vector(const DeviceCodebook &cb, const float *projection, const void *codes,
                                    int code_width, ptrdiff_t cstride, float *out, ptrdiff_t ostride, float *scratch,
                                    int *err_flag, cudaStream_t stream)
{
    reconstruct_vector_kernel<<<1, kVecThreads, 0, stream>>>(cb.quantizers, (int)cb.M, (int)cb.k, (int)cb.dsub,
                                                             projection, codes, code_width, (long long)cstride, out,
                                                             (long long)ostride, scratch, err_flag);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

}  // namespace rb
