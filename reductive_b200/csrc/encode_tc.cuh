// encode_tc.cuh — interface of the tcgen05 tensor-core encode path (see encode_tc.cu).
#pragma once
#include "common.cuh"

namespace rb {

// Codebook operands in the layout the tensor kernel consumes (built once per codebook).
struct TensorOperands {
    void *b_tiles = nullptr;   // bf16 split codebook, UMMA K-major core-matrix layout, per subquantizer
    float *row_bound = nullptr;  // [M] per-subquantizer max ||c||, for the error bound
    size_t bytes = 0;
    int kpad = 0;       // K extent (multiple of 16) of the augmented operand
    bool ready() const { return b_tiles != nullptr; }
    rb_status prepare(const DeviceCodebook &cb, cudaStream_t stream);
    void release();
    void release_async(cudaStream_t stream);
};

bool tensor_path_supported(const DeviceCodebook &cb);

rb_status launch_encode_tensor(const DeviceCodebook &cb, const TensorOperands &tc, const float *x, size_t n,
                               ptrdiff_t ldx, void *codes, int code_width, ptrdiff_t crs, ptrdiff_t ccs,
                               cudaStream_t stream);

}  // namespace rb
