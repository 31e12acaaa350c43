// Semantics probe for cp.async.bulk.tensor.2d ... tile::gather4 on sm_100a: which box shape the tensor map needs and
// where the four gathered rows land in shared memory.   nvcc -gencode arch=compute_100a,code=sm_100a -o gather4_probe gather4_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

__global__ void k(const __grid_constant__ CUtensorMap tm, float *out, int col, int r0, int r1, int r2, int r3, int bytes)
{
    __shared__ __align__(128) float buf[64];
    __shared__ uint64_t bar;
    const unsigned sb = (unsigned)__cvta_generic_to_shared(&bar), sd = (unsigned)__cvta_generic_to_shared(buf);
    if (threadIdx.x < 64) buf[threadIdx.x] = -1.f;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sb));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb), "r"(bytes));
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(sd),
            "l"(&tm), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(sb)
            : "memory");
    }
    __syncthreads();
    unsigned ok = 0;
    for (int spin = 0; spin < (1 << 22) && !ok; spin++)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.b32 %0,1,0,p; }" : "=r"(ok) : "r"(sb));
    if (threadIdx.x < 64) out[threadIdx.x] = buf[threadIdx.x];
    if (threadIdx.x == 0) out[64] = ok ? 1.f : 0.f;
}

int main()
{
    const int rows = 256, d = 64;
    std::vector<float> h(rows * d);
    for (int r = 0; r < rows; r++)
        for (int c = 0; c < d; c++) h[r * d + c] = r * 100.f + c;
    float *x, *out;
    cudaMalloc(&x, h.size() * 4);
    cudaMalloc(&out, 65 * 4);
    cudaMemcpy(x, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeFn enc = (EncodeFn)fn;
    for (int boxrows : {1, 4}) {
        CUtensorMap tm;
        const cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)rows};
        const cuuint64_t strides[1] = {(cuuint64_t)d * 4};
        const cuuint32_t box[2] = {8, (cuuint32_t)boxrows};
        const cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("box rows %d: encode -> %d\n", boxrows, (int)r);
        if (r != CUDA_SUCCESS) continue;
        cudaMemset(out, 0, 65 * 4);
        k<<<1, 64>>>(tm, out, 16, 5, 17, 3, 99, 128);
        cudaError_t e = cudaDeviceSynchronize();
        printf("  kernel: %s\n", cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        float o[65];
        cudaMemcpy(o, out, 65 * 4, cudaMemcpyDeviceToHost);
        printf("  completed=%g  smem:", o[64]);
        for (int i = 0; i < 40; i++) printf(" %g", o[i]);
        printf("\n");
    }
    return 0;
}
