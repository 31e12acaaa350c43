// encode_tc.cuh — interface of the tcgen05 tensor-core encode path (see encode_tc.cu).
#pragma once
#include "common.cuh"

namespace rb {

// Codebook operands in the layout the tensor kernel consumes (rebuilt whenever the centroids change).
struct TensorOperands {
    void *b_tiles = nullptr;     // fp16 two-limb codebook, UMMA K-major core-matrix layout: [M][kpad/8][256][8 halves]
    float *consts = nullptr;     // [M] max_j ||c_j||^2 per subquantizer, then 4 floats: scale 2^e, 2^2e, bad flag, |c|max
    size_t bytes = 0;
    int kpad = 0;                // K extent (multiple of 16) of the augmented operand: 3*dsub + 2 rounded up
    bool ready() const { return b_tiles != nullptr; }
    rb_status prepare(const DeviceCodebook &cb, cudaStream_t stream);
    void release();
    void release_async(cudaStream_t stream);
};

// Shapes the tensor kernel covers (k == 256 centroids, subvector widths it is instantiated for).
bool tensor_path_supported(const DeviceCodebook &cb);
// ... and the per-call conditions (alignment of x / ldx, n < 2^32).
bool tensor_call_supported(const DeviceCodebook &cb, const float *x, size_t n, ptrdiff_t ldx);

// The batch handed to the encode is an approximate rotation x ~ x0 . r (project_tc.cu) held in a buffer the caller
// owns: every component of row i may be off by rowerr[i] + err_floor / *sx_dev.  The kernel widens its margin by
// that and the flagged (row, subquantizer) pairs are re-rotated exactly before they are re-decided.
struct RotatedInput {
    const float *x0;      // the un-rotated rows
    ptrdiff_t ldx0;
    const float *r;       // row-major [d][d]
    size_t d;
    const float *rowerr;  // per row; NaN: the row could not be rotated approximately
    const float *sx_dev;  // device scalar: the power-of-two operand scale the rotation used
    float err_floor;
};

rb_status launch_encode_tensor(const DeviceCodebook &cb, const TensorOperands &tc, const float *x, size_t n,
                               ptrdiff_t ldx, void *codes, int code_width, ptrdiff_t crs, ptrdiff_t ccs,
                               cudaStream_t stream, const RotatedInput *rot = nullptr);

}  // namespace rb
