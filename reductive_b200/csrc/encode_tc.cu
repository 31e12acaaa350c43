// encode_tc.cu — tcgen05 tensor-core encode path (placeholder until the kernel lands: reports "not ready",
// so RB_ENCODE_AUTO resolves to the exact SIMT kernel and RB_ENCODE_TENSOR fails loudly).
#include "encode_tc.cuh"

namespace rb {

bool tensor_path_supported(const DeviceCodebook &) { return false; }
rb_status TensorOperands::prepare(const DeviceCodebook &, cudaStream_t) { return RB_OK; }
void TensorOperands::release() {}
void TensorOperands::release_async(cudaStream_t) {}

rb_status launch_encode_tensor(const DeviceCodebook &, const TensorOperands &, const float *, size_t, ptrdiff_t,
                               void *, int, ptrdiff_t, ptrdiff_t, cudaStream_t)
{
    set_error("tensor encode path not built");
    return RB_ERR_UNSUPPORTED;
}

}  // namespace rb
