// project_tc.cuh — interface of the tcgen05 rotation GEMM (see project_tc.cu).
#pragma once
#include "common.cuh"

namespace rb {

// R split into scaled FP16 limbs in the layout the kernel streams (built once per quantizer handle).
struct ProjTensorOperands {
    void *bop = nullptr;  // [n_groups][n_chunks][2 limbs][4 core columns][NT][8 halves]
    size_t bytes = 0;
    int d = 0, NT = 0, n_groups = 0, n_chunks = 0;
    float sr = 1.f;       // power-of-two scale applied to R
    float rabsmax = 0.f;  // max |r_ij|
    float rcolmax = 0.f;  // max_j ||R[:, j]||_2 (rounded up)
    float *chunk_w = nullptr;  // device [n_chunks]: error weight of each K chunk (accumulation-order error model)
    float limb_coef = 0.f;     // operand-split error per unit of ||x||
    bool ready() const { return bop != nullptr; }
    // r_dev / r_host: the same row-major [d][d] matrix (y = x . r); leaves the operands empty when the shape is
    // not covered or r holds non-finite values
    rb_status prepare(const float *r_dev, const float *r_host, size_t d, cudaStream_t stream);
    void release();
};

bool project_tensor_shape_supported(size_t d);
bool project_tensor_call_supported(const ProjTensorOperands &ops, const float *x, size_t n, ptrdiff_t ldx, const float *y,
                                   ptrdiff_t ldy);
// ... and whether the kernel can also produce the per-row error bound the encode needs (d <= 1024)
bool project_tensor_rowerr_supported(const ProjTensorOperands &ops);
// power-of-two scale for x from a known bound on |x|
float project_scale_for_absmax(float amax);
// ... or from a strided sample of rows, on the device: scratch4[0] receives the scale (scratch4 = 4 floats)
rb_status launch_project_sample_scale(const float *x, size_t n, size_t d, ptrdiff_t ldx, float *scratch4, cudaStream_t stream);
// absolute part of the error bound, in units of 1 / (operand scale of x)
float project_tensor_error_floor(const ProjTensorOperands &ops);
// y~ = x . R; rowerr (optional, n floats) receives a bound on |y~_j - y_j| for every component of the row against
// the reference's FP32 GEMM (without the 1 / scale part above), NaN for rows the split cannot represent
// (non-finite values, |x * scale| >= 2^15) whose outputs are then unspecified
rb_status launch_project_tensor(const ProjTensorOperands &ops, const float *x, size_t n, ptrdiff_t ldx, const float *sx_dev,
                                float sx_host, float *y, ptrdiff_t ldy, float *rowerr, cudaStream_t stream);

}  // namespace rb
