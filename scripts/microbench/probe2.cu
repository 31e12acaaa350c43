// probe2.cu — round-2 micro-measurements for the packed-16-bit epilogue of the tensor encode kernel:
//   1. tcgen05.mma kind::f16 with an F16 accumulator (D format 0): where the values land in tensor memory
//      (one per 32-bit cell or two), what tcgen05.ld.pack::16b returns, rounding of the final conversion and of
//      a second accumulating instruction;
//   2. issue rates per SM of HMNMX2 (min.f16x2), VIMNMX3.S16x2, the f16 -> f32 conversion, FFMA, FFMA2 alone and in
//      pairs on independent registers (does a pair of pipes overlap?).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o probe2 probe2.cu
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

#include "../../reductive_b200/csrc/sm100_ptx.cuh"

using namespace rb::ptx;

#define CK(x)                                                                                      \
    do {                                                                                           \
        cudaError_t e = (x);                                                                       \
        if (e != cudaSuccess) {                                                                    \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);        \
            exit(1);                                                                               \
        }                                                                                          \
    } while (0)

// A: [128, K] fp16, B: [256, K] fp16 row-major in global memory.  D format f16.
__global__ void __launch_bounds__(128) mma_f16d_probe(const __half *A, const __half *B, int K, uint32_t *Draw, uint32_t *Dpack)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int nchunk = K / 8;
    unsigned char *sA = smem;
    unsigned char *sB = smem + (size_t)nchunk * 128 * 16;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < nchunk * 128; i += 128) {
        const int c = i / 128, r = i % 128;
        *reinterpret_cast<uint4 *>(sA + (size_t)i * 16) = *reinterpret_cast<const uint4 *>(A + (size_t)r * K + c * 8);
    }
    for (int i = tid; i < nchunk * 256; i += 128) {
        const int c = i / 256, r = i % 256;
        *reinterpret_cast<uint4 *>(sB + (size_t)i * 16) = *reinterpret_cast<const uint4 *>(B + (size_t)r * K + c * 8);
    }
    fence_proxy_async_smem();
    if (warp == 0) {
        tmem_alloc(&tmem_slot, 512);
        tmem_relinquish();
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_slot;
    // poison the 512 columns so untouched cells are recognisable
    {
        const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < 512; c0 += 8) {
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(lane_base + c0),
                         "r"(0xdeadbeefu)
                         : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // D format f16: bits [4,6) = 0
    const uint32_t idesc = (0u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    if (warp == 0) {
        if (elect_one()) {
            for (int ks = 0; ks < K / 16; ks++) {
                const uint64_t ad = smem_desc_kmajor(smem_u32(sA) + ks * 2 * 128 * 16, 128 * 16, 128);
                const uint64_t bd = smem_desc_kmajor(smem_u32(sB) + ks * 2 * 256 * 16, 256 * 16, 128);
                mma_f16_ss(tbase, ad, bd, idesc, ks > 0 ? 1u : 0u);
            }
            tc_commit(&bar);
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    const int row = warp * 32 + (tid & 31);
    for (int c0 = 0; c0 < 512; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; j++) Draw[(size_t)row * 512 + c0 + j] = v[j];
    }
    for (int c0 = 0; c0 < 512; c0 += 64) {
        uint32_t v[32];
        tmem_ld32_pack16(tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; j++) Dpack[(size_t)row * 256 + c0 / 2 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

// ---------------------------------------------------------------------------------------------------------
// issue rates.  8 independent chains per thread.
// ---------------------------------------------------------------------------------------------------------
enum Op { FMNMX3, VIMNMX3_16, VIMNMX2_16, HMNMX2, CVT_H2F, FFMA, FFMA2, FMASAT, MIX_H_V16, MIX_H_FFMA, MIX_V16_FFMA, MIX_V16_FFMA2,
          MIX_V16_FMASAT, MIX_H_FMNMX3, MIX_V16_CVT, MIX_V16_2FFMA, NOPS };
static const char *op_names[] = {"FMNMX3", "VIMNMX3.S16x2", "VIMNMX.S16x2(2-in)", "HMNMX2", "cvt f16x2->2xf32", "FFMA", "FFMA2", "FFMA.SAT",
                                 "HMNMX2 + VIMNMX3.S16x2", "HMNMX2 + FFMA", "VIMNMX3.S16x2 + FFMA", "VIMNMX3.S16x2 + FFMA2",
                                 "VIMNMX3.S16x2 + FFMA.SAT", "HMNMX2 + FMNMX3", "VIMNMX3.S16x2 + cvt", "VIMNMX3.S16x2 + 2 FFMA"};
static const int op_count[] = {1, 1, 1, 1, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 3, 3};

template <int OP>
__global__ void __launch_bounds__(512) alu_bench(int iters, unsigned seed, unsigned *sink, long long *cycles)
{
    unsigned r[8], x[8], y[8], s[8];
    unsigned long long w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r[i] = seed * (threadIdx.x + 1) + i * 0x9e3779b9u;
        x[i] = (r[i] ^ 0x12345678u) & 0x3fff3fffu;
        y[i] = (r[i] * 3u + 1u) & 0x3fff3fffu;
        s[i] = 0x3f800000u + i;
        w[i] = ((unsigned long long)s[i] << 32) | s[i];
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (OP == FMNMX3) {
                    asm volatile("min.f32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(x[i]), "r"(y[i]));
                } else if (OP == VIMNMX3_16) {
                    r[i] = __vimin3_s16x2(r[i], x[i], y[i]);
                    asm volatile("" : "+r"(r[i]));
                } else if (OP == VIMNMX2_16) {
                    r[i] = __vmins2(r[i], x[i]);
                    asm volatile("" : "+r"(r[i]));
                } else if (OP == HMNMX2) {
                    asm volatile("min.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(x[i]));
                } else if (OP == CVT_H2F) {
                    // two conversions per step (low and high half)
                    unsigned a, b;
                    asm volatile("{.reg .f16 lo, hi; mov.b32 {lo, hi}, %2; cvt.f32.f16 %0, lo; cvt.f32.f16 %1, hi;}" : "=r"(a), "=r"(b) : "r"(r[i]));
                    r[i] = a ^ b;  // keeps a dependency (LOP3 on the ALU pipe; counted as noise)
                } else if (OP == FFMA) {
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(s[i]) : "r"(x[i]), "r"(y[i]));
                } else if (OP == FFMA2) {
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(w[i]) : "l"(((unsigned long long)x[i] << 32) | y[i]));
                } else if (OP == FMASAT) {
                    asm volatile("fma.rn.sat.f32 %0, %0, %1, %2;" : "+r"(s[i]) : "r"(x[i]), "r"(y[i]));
                } else if (OP == MIX_H_V16) {
                    asm volatile("min.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(x[i]));
                    s[i] = __vimin3_s16x2(s[i], x[i], y[i]);
                    asm volatile("" : "+r"(s[i]));
                } else if (OP == MIX_H_FFMA) {
                    asm volatile("min.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(x[i]));
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(s[i]) : "r"(x[i]), "r"(y[i]));
                } else if (OP == MIX_V16_FFMA) {
                    r[i] = __vimin3_s16x2(r[i], x[i], y[i]);
                    asm volatile("" : "+r"(r[i]));
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(s[i]) : "r"(x[i]), "r"(y[i]));
                } else if (OP == MIX_V16_2FFMA) {
                    r[i] = __vimin3_s16x2(r[i], x[i], y[i]);
                    asm volatile("" : "+r"(r[i]));
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(s[i]) : "r"(x[i]), "r"(y[i]));
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(s[(i + 1) & 7]), "r"(y[i]));
                } else if (OP == MIX_V16_FFMA2) {
                    r[i] = __vimin3_s16x2(r[i], x[i], y[i]);
                    asm volatile("" : "+r"(r[i]));
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(w[i]) : "l"(((unsigned long long)x[i] << 32) | y[i]));
                } else if (OP == MIX_V16_FMASAT) {
                    r[i] = __vimin3_s16x2(r[i], x[i], y[i]);
                    asm volatile("" : "+r"(r[i]));
                    asm volatile("fma.rn.sat.f32 %0, %0, %1, %2;" : "+r"(s[i]) : "r"(x[i]), "r"(y[i]));
                } else if (OP == MIX_H_FMNMX3) {
                    asm volatile("min.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(x[i]));
                    asm volatile("min.f32 %0, %0, %1, %2;" : "+r"(s[i]) : "r"(x[i]), "r"(y[i]));
                } else if (OP == MIX_V16_CVT) {
                    r[i] = __vimin3_s16x2(r[i], x[i], y[i]);
                    asm volatile("" : "+r"(r[i]));
                    unsigned a, b;
                    asm volatile("{.reg .f16 lo, hi; mov.b32 {lo, hi}, %2; cvt.f32.f16 %0, lo; cvt.f32.f16 %1, hi;}" : "=r"(a), "=r"(b) : "r"(x[i]));
                    asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+r"(s[i]) : "r"(a), "r"(b));
                }
            }
        }
    }
    const long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= r[i] ^ s[i] ^ (unsigned)w[i] ^ (unsigned)(w[i] >> 32) ^ x[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run_alu(unsigned *sink, long long *cyc_d)
{
    const int iters = 2000, blocks = 148 * 2, threads = 512;
    alu_bench<OP><<<blocks, threads>>>(10, 1, sink, cyc_d);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    alu_bench<OP><<<blocks, threads>>>(iters, 7, sink, cyc_d);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> cyc(blocks);
    CK(cudaMemcpy(cyc.data(), cyc_d, blocks * sizeof(long long), cudaMemcpyDeviceToHost));
    double mean = 0;
    for (auto c : cyc) mean += (double)c;
    mean /= blocks;
    const double steps = 2.0 * 512 * (double)iters * 64;
    printf("ALU %-26s  %7.1f steps/clk/SM  %7.1f counted-instr/clk/SM   (%.3f ms, clk rate %.3f GHz)\n", op_names[OP], steps / mean,
           steps * op_count[OP] / mean, ms, mean / (ms * 1e6));
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device: %s, %d SMs, cc %d.%d\n", prop.name, prop.multiProcessorCount, prop.major, prop.minor);

    for (int K : {16, 32}) {
        std::vector<__half> hA(128 * K), hB(256 * K);
        std::vector<double> dA(128 * K), dB(256 * K);
        srand(99 + K);
        for (int i = 0; i < 128 * K; i++) {
            const double v = ((rand() % 8192) - 4096) / 1024.0;  // [-4, 4)
            hA[i] = __float2half((float)v);
            dA[i] = (double)__half2float(hA[i]);
        }
        for (int i = 0; i < 256 * K; i++) {
            const double v = ((rand() % 8192) - 4096) / 2048.0;  // [-2, 2)
            hB[i] = __float2half((float)v);
            dB[i] = (double)__half2float(hB[i]);
        }
        __half *A, *B;
        uint32_t *Draw, *Dp;
        CK(cudaMalloc(&A, hA.size() * 2));
        CK(cudaMalloc(&B, hB.size() * 2));
        CK(cudaMalloc(&Draw, 128 * 512 * 4));
        CK(cudaMalloc(&Dp, 128 * 256 * 4));
        CK(cudaMemcpy(A, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(B, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
        const size_t smem = (size_t)(K / 8) * (128 + 256) * 16;
        mma_f16d_probe<<<1, 128, smem>>>(A, B, K, Draw, Dp);
        CK(cudaDeviceSynchronize());
        std::vector<uint32_t> raw(128 * 512), pk(128 * 256);
        CK(cudaMemcpy(raw.data(), Draw, raw.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(pk.data(), Dp, pk.size() * 4, cudaMemcpyDeviceToHost));
        // which cells were written?
        int written_cols = 0;
        for (int c = 0; c < 512; c++) {
            bool any = false;
            for (int r = 0; r < 128; r++) any |= raw[r * 512 + c] != 0xdeadbeefu;
            written_cols += any;
        }
        printf("F16 accumulator, K=%d: %d of 512 TMEM columns written\n", K, written_cols);
        printf("  row 3 raw cells 0..7: ");
        for (int c = 0; c < 8; c++) printf("%08x ", raw[3 * 512 + c]);
        printf("\n  row 3 packed regs 0..3: ");
        for (int c = 0; c < 4; c++) printf("%08x ", pk[3 * 256 + c]);
        printf("\n");
        // hypotheses: (a) cell c low 16 bits = f16(D[r][c]);  (b) cell c = f16 pair (D[r][2c], D[r][2c+1])
        int ok_a = 0, ok_b = 0, rn = 0, rz = 0, tot = 0;
        double max_ulp = 0;
        for (int r = 0; r < 128; r++)
            for (int c = 0; c < 256; c++) {
                double ref = 0;
                for (int k = 0; k < K; k++) ref += dA[r * K + k] * dB[c * K + k];
                const __half h_rn = __float2half_rn((float)ref);  // (double -> float is exact enough here: values are dyadic)
                const __half h_rz = __float2half_rz((float)ref);
                const uint16_t a = (uint16_t)(raw[r * 512 + c] & 0xffffu);
                const uint32_t cell_b = raw[r * 512 + c / 2];
                const uint16_t b = (uint16_t)((c & 1) ? (cell_b >> 16) : (cell_b & 0xffffu));
                const uint16_t want_rn = *reinterpret_cast<const uint16_t *>(&h_rn);
                const uint16_t want_rz = *reinterpret_cast<const uint16_t *>(&h_rz);
                if (a == want_rn || a == want_rz) ok_a++;
                if (b == want_rn || b == want_rz) ok_b++;
                const uint16_t got = written_cols > 128 ? a : b;
                if (got == want_rn) rn++;
                if (got == want_rz) rz++;
                __half gh;
                memcpy(&gh, &got, 2);
                const double g = (double)__half2float(gh);
                const double ulp = ldexp(1.0, (int)floor(log2(fmax(fabs(ref), 1e-30))) - 10);
                max_ulp = fmax(max_ulp, fabs(g - ref) / ulp);
                tot++;
            }
        printf("  layout (a: one f16 per cell) matches %d, (b: two per cell) matches %d of %d;  == RN %d, == RZ %d, max err %.3f ulp_f16\n",
               ok_a, ok_b, tot, rn, rz, max_ulp);
        // pack::16b
        int pk_a = 0;
        for (int r = 0; r < 128; r++)
            for (int j = 0; j < 128; j++) {
                const uint32_t want = (raw[r * 512 + 2 * j] & 0xffffu) | (raw[r * 512 + 2 * j + 1] << 16);
                if (pk[r * 256 + j] == want) pk_a++;
            }
        printf("  pack::16b reg j == lo16(cell 2j) | lo16(cell 2j+1) << 16 for %d of %d\n", pk_a, 128 * 128);
        cudaFree(A); cudaFree(B); cudaFree(Draw); cudaFree(Dp);
    }

    unsigned *sink2;
    long long *cyc2;
    CK(cudaMalloc(&sink2, 148 * 2 * 512 * 4));
    CK(cudaMalloc(&cyc2, 148 * 2 * 8));
    run_alu<FMNMX3>(sink2, cyc2);
    run_alu<VIMNMX3_16>(sink2, cyc2);
    run_alu<VIMNMX2_16>(sink2, cyc2);
    run_alu<HMNMX2>(sink2, cyc2);
    run_alu<CVT_H2F>(sink2, cyc2);
    run_alu<FFMA>(sink2, cyc2);
    run_alu<FFMA2>(sink2, cyc2);
    run_alu<FMASAT>(sink2, cyc2);
    run_alu<MIX_H_V16>(sink2, cyc2);
    run_alu<MIX_H_FFMA>(sink2, cyc2);
    run_alu<MIX_V16_FFMA>(sink2, cyc2);
    run_alu<MIX_V16_2FFMA>(sink2, cyc2);
    run_alu<MIX_V16_FFMA2>(sink2, cyc2);
    run_alu<MIX_V16_FMASAT>(sink2, cyc2);
    run_alu<MIX_H_FMNMX3>(sink2, cyc2);
    run_alu<MIX_V16_CVT>(sink2, cyc2);
    printf("probe2 done\n");
    return 0;
}
