// probe.cu — B200 micro-measurements that the tensor encode kernel's design rests on (run under gpurun):
//   1. tcgen05.mma kind::f16 correctness with the no-swizzle K-major layout the kernel uses, accumulation
//      behaviour with a large "magic" addend (fixed-point trick) and the semantics of tcgen05.ld.pack::16b;
//   2. cycles per tcgen05.mma (M=128, N=256, K=16);
//   3. tcgen05.ld throughput per SM (4 and 8 warps);
//   4. issue rate per SM of the candidate epilogue instructions (FMNMX3, VIMNMX3, VIMNMX3.S16x2, LOP3, LEA, IMAD,
//      FADD, VIADDMNMX) alone and in pairs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o probe probe.cu
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

// ---------------------------------------------------------------------------------------------------------
// 1. MMA correctness.  A: [128, K] fp16, B: [256, K] fp16 given row-major in global memory; K = 16 * ksteps.
//    shared layout: [k_chunk (K/8)][row][8 halves]  (core matrix = 8 rows x 16 B contiguous)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) mma_probe(const __half *A, const __half *B, int K, float *D, uint32_t *Dpack,
                                                 long long *cycles, int reps)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int nchunk = K / 8;
    unsigned char *sA = smem;                       // nchunk * 128 * 16
    unsigned char *sB = smem + (size_t)nchunk * 128 * 16;  // nchunk * 256 * 16
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
    const uint32_t idesc = idesc_f16(128, 256, 0);
    uint32_t parity = 0;
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
        if (elect_one()) {
            t0 = clock64();
            for (int rep = 0; rep < reps; rep++) {
                for (int ks = 0; ks < K / 16; ks++) {
                    const uint64_t ad = smem_desc_kmajor(smem_u32(sA) + ks * 2 * 128 * 16, 128 * 16, 128);
                    const uint64_t bd = smem_desc_kmajor(smem_u32(sB) + ks * 2 * 256 * 16, 256 * 16, 128);
                    mma_f16_ss(tbase + (rep & 1) * 256, ad, bd, idesc, ks > 0 ? 1u : 0u);
                }
            }
            tc_commit(&bar);
        }
        __syncwarp();
    }
    mbar_wait(&bar, parity);
    if (tid == 0) {
        t1 = clock64();
        cycles[0] = t1 - t0;
    }
    tc_fence_after();
    // read back accumulator 0 (or 1 when reps is even -> last written is (reps-1)&1)
    const uint32_t acc = tbase + ((reps - 1) & 1) * 256;
    const int row = warp * 32 + (tid & 31);
    for (int c0 = 0; c0 < 256; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(acc + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; j++) D[(size_t)row * 256 + c0 + j] = __uint_as_float(v[j]);
    }
    for (int c0 = 0; c0 < 256; c0 += 64) {
        uint32_t v[32];
        tmem_ld32_pack16(acc + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; j++) Dpack[(size_t)row * 128 + c0 / 2 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

// ---------------------------------------------------------------------------------------------------------
// 3. tcgen05.ld throughput.  Each warp reads its 32 lanes x 256 columns `iters` times.
// ---------------------------------------------------------------------------------------------------------
template <int MODE>  // 0: x32, 1: x32 pack16 (64 columns per instruction), 2: x16
__global__ void __launch_bounds__(256) ldtm_bench(int iters, unsigned *sink, long long *cycles)
{
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        tmem_alloc(&tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
    unsigned acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int c0 = 0; c0 < 256; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tbase + c0 + ((warp >> 2) * 256), v);
                tmem_wait_ld();
                acc ^= v[0] ^ v[31];
            }
        } else if (MODE == 1) {
#pragma unroll
            for (int c0 = 0; c0 < 256; c0 += 64) {
                uint32_t v[32];
                tmem_ld32_pack16(tbase + c0 + ((warp >> 2) * 256), v);
                tmem_wait_ld();
                acc ^= v[0] ^ v[31];
            }
        } else {
#pragma unroll
            for (int c0 = 0; c0 < 256; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(tbase + c0 + ((warp >> 2) * 256), v);
                tmem_wait_ld();
                acc ^= v[0] ^ v[15];
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + tid] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_slot, 512);
}

// ---------------------------------------------------------------------------------------------------------
// 4. ALU issue rates.  8 independent chains per thread, 8 ops per chain per iteration.
// ---------------------------------------------------------------------------------------------------------
enum Op { FMNMX3, VIMNMX3, VIMNMX3_16, LOP3, LEA, IMAD, FADD, VIADDMNMX, FMNMX2, MIX_IMAD_VIMNMX3, MIX_FADD_FMNMX3,
          MIX_LOP3_VIMNMX3, MIX_LEA_VIMNMX3, PRMT, FFMA, LEA_IMM, LOP3_IMM, MIX_LEAIMM_VIMNMX3, MIX_LOP3IMM_VIMNMX3, MIX_LOP3IMM_FMNMX3, FADD_IMM, MIX_FADDIMM_FMNMX3, NOPS };
static const char *op_names[] = {"FMNMX3", "VIMNMX3.S32", "VIMNMX3.S16x2", "LOP3", "LEA(shl+add)", "IMAD", "FADD",
                                 "VIADDMNMX", "FMNMX(2-in)", "IMAD+0.5*VIMNMX3", "FADD+0.5*FMNMX3", "LOP3+0.5*VIMNMX3",
                                 "LEA+0.5*VIMNMX3", "PRMT", "FFMA", "LEA imm,imm", "LOP3 R,R,imm", "2 LEAimm + VIMNMX3", "2 LOP3imm + VIMNMX3", "2 LOP3imm + FMNMX3", "FADD R,imm", "2 FADDimm + FMNMX3"};
// instructions counted per inner step (for the mixes: 3 = 2 packs + 1 min3)
static const int op_count[] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 3, 3, 3, 3, 1, 1, 1, 1, 3, 3, 3, 1, 3};

template <int OP>
__global__ void __launch_bounds__(512) alu_bench(int iters, unsigned seed, unsigned *sink, long long *cycles)
{
    unsigned r[8], x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r[i] = seed * (threadIdx.x + 1) + i * 0x9e3779b9u;
        x[i] = r[i] ^ 0x12345678u;
        y[i] = r[i] * 3u + 1u;
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
                } else if (OP == FMNMX2) {
                    asm volatile("min.f32 %0, %0, %1;" : "+r"(r[i]) : "r"(x[i]));
                } else if (OP == VIMNMX3) {
                    r[i] = __vimin3_s32(r[i], x[i], y[i]);
                    asm volatile("" : "+r"(r[i]));
                } else if (OP == VIMNMX3_16) {
                    r[i] = __vimin3_s16x2(r[i], x[i], y[i]);
                    asm volatile("" : "+r"(r[i]));
                } else if (OP == LOP3) {
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(r[i]) : "r"(x[i]), "r"(y[i]));  // (a&b)|c
                } else if (OP == LEA) {
                    r[i] = r[i] * 256u + x[i];
                    asm volatile("" : "+r"(r[i]));
                } else if (OP == IMAD) {
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(x[i]), "r"(y[i]));
                } else if (OP == FADD) {
                    asm volatile("add.rn.f32 %0, %0, %1;" : "+r"(r[i]) : "r"(x[i]));
                } else if (OP == FFMA) {
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(x[i]), "r"(y[i]));
                } else if (OP == VIADDMNMX) {
                    r[i] = __viaddmin_s32(x[i], y[i], r[i]);
                    asm volatile("" : "+r"(r[i]));
                } else if (OP == PRMT) {
                    asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(x[i]), "r"(y[i]));
                } else if (OP == MIX_IMAD_VIMNMX3) {
                    unsigned a, b;
                    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a) : "r"(x[i]), "r"(y[(i + 1) & 7]), "r"(r[(i + 1) & 7]));
                    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(b) : "r"(y[i]), "r"(x[(i + 1) & 7]), "r"(r[(i + 2) & 7]));
                    r[i] = __vimin3_u32(r[i], a, b);
                    asm volatile("" : "+r"(r[i]));
                } else if (OP == MIX_FADD_FMNMX3) {
                    unsigned a, b;
                    asm volatile("add.rn.f32 %0, %1, %2;" : "=r"(a) : "r"(x[i]), "r"(r[(i + 1) & 7]));
                    asm volatile("add.rn.f32 %0, %1, %2;" : "=r"(b) : "r"(y[i]), "r"(r[(i + 2) & 7]));
                    asm volatile("min.f32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
                } else if (OP == MIX_LOP3_VIMNMX3) {
                    unsigned a, b;
                    asm volatile("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(a) : "r"(x[i]), "r"(y[(i + 1) & 7]), "r"(r[(i + 1) & 7]));
                    asm volatile("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(b) : "r"(y[i]), "r"(x[(i + 1) & 7]), "r"(r[(i + 2) & 7]));
                    r[i] = __vimin3_u32(r[i], a, b);
                    asm volatile("" : "+r"(r[i]));
                } else if (OP == LEA_IMM) {
                    r[i] = r[i] * 256u + (unsigned)(u * 8 + i);
                    asm volatile("" : "+r"(r[i]));
                } else if (OP == LOP3_IMM) {
                    r[i] = (r[i] & x[0]) | (unsigned)(u * 8 + i + 1);
                    asm volatile("" : "+r"(r[i]));
                } else if (OP == FADD_IMM) {
                    asm volatile("add.rn.f32 %0, %0, %1;" : "+r"(r[i]) : "r"(0x3f800000u + u * 8 + i));
                } else if (OP == MIX_LEAIMM_VIMNMX3) {
                    unsigned a = x[i] * 256u + (unsigned)(u * 8 + i);
                    unsigned b = y[i] * 256u + (unsigned)(u * 8 + i + 64);
                    asm volatile("" : "+r"(a), "+r"(b));
                    r[i] = __vimin3_u32(r[i], a, b);
                    asm volatile("" : "+r"(r[i]));
                    x[i] ^= r[i];
                } else if (OP == MIX_LOP3IMM_VIMNMX3) {
                    unsigned a = (x[i] & y[7]) | (unsigned)(u * 8 + i);
                    unsigned b = (y[i] & y[7]) | (unsigned)(u * 8 + i + 64);
                    asm volatile("" : "+r"(a), "+r"(b));
                    r[i] = __vimin3_u32(r[i], a, b);
                    asm volatile("" : "+r"(r[i]));
                } else if (OP == MIX_LOP3IMM_FMNMX3) {
                    unsigned a = (x[i] & y[7]) | (unsigned)(u * 8 + i);
                    unsigned b = (y[i] & y[7]) | (unsigned)(u * 8 + i + 64);
                    asm volatile("" : "+r"(a), "+r"(b));
                    asm volatile("min.f32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
                } else if (OP == MIX_FADDIMM_FMNMX3) {
                    unsigned a, b;
                    asm volatile("add.rn.f32 %0, %1, %2;" : "=r"(a) : "r"(x[i]), "r"(0x3f800000u + u * 8 + i));
                    asm volatile("add.rn.f32 %0, %1, %2;" : "=r"(b) : "r"(y[i]), "r"(0x3f800000u + u * 8 + i + 64));
                    asm volatile("min.f32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
                } else if (OP == MIX_LEA_VIMNMX3) {
                    unsigned a = r[(i + 1) & 7] * 256u + x[i];
                    unsigned b = r[(i + 2) & 7] * 256u + y[i];
                    asm volatile("" : "+r"(a), "+r"(b));
                    r[i] = __vimin3_u32(r[i], a, b);
                    asm volatile("" : "+r"(r[i]));
                }
            }
        }
    }
    const long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= r[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run_alu(unsigned *sink, long long *cyc_d)
{
    const int iters = 2000, blocks = 148 * 2, threads = 512;  // 32 warps / SM resident (2 blocks x 16 warps)
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
    // per SM: 2 blocks x 512 threads x iters x 64 steps x op_count instructions (thread-level)
    const double thread_instr = 2.0 * 512 * (double)iters * 64 * op_count[OP];
    printf("ALU %-20s  %8.1f thread-instr/clk64/SM   (%.0f clk64, %.3f ms -> clk64 rate %.3f GHz; %.1f thread-instr/ns/SM)\n", op_names[OP], thread_instr / mean, mean, ms, mean / (ms * 1e6), thread_instr / (ms * 1e6));
}

int main(int argc, char **argv)
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device: %s, %d SMs, cc %d.%d\n", prop.name, prop.multiProcessorCount, prop.major, prop.minor);

    // ---- 1/2: MMA probe ------------------------------------------------------------------------------------
    for (int test = 0; test < 3; test++) {
        const int K = 32;
        std::vector<__half> hA(128 * K), hB(256 * K);
        std::vector<double> dA(128 * K), dB(256 * K);
        srand(1234 + test);
        const double magic = test == 1 ? 49152.0 : (test == 2 ? 384.0 : 0.0);  // 1.5*2^15 (ulp 2^-8), 1.5*2^8 (ulp 2^-15)
        for (int r = 0; r < 128; r++)
            for (int k = 0; k < K; k++) {
                double v;
                if (test == 0) v = (rand() % 17) - 8;  // small integers: exact
                else v = ((rand() % 4096) - 2048) / 512.0;  // multiples of 2^-9 in [-4, 4)
                if (test > 0 && k == K - 1) v = 1.0;
                hA[r * K + k] = __float2half((float)v);
                dA[r * K + k] = (double)__half2float(hA[r * K + k]);
            }
        for (int r = 0; r < 256; r++)
            for (int k = 0; k < K; k++) {
                double v;
                if (test == 0) v = (rand() % 9) - 4;
                else v = ((rand() % 4096) - 2048) / 1024.0;  // multiples of 2^-10 in [-2, 2)
                if (test > 0 && k == K - 1) v = magic;
                hB[r * K + k] = __float2half((float)v);
                dB[r * K + k] = (double)__half2float(hB[r * K + k]);
            }
        __half *A, *B;
        float *D;
        uint32_t *Dp;
        long long *cyc;
        CK(cudaMalloc(&A, hA.size() * 2));
        CK(cudaMalloc(&B, hB.size() * 2));
        CK(cudaMalloc(&D, 128 * 256 * 4));
        CK(cudaMalloc(&Dp, 128 * 128 * 4));
        CK(cudaMalloc(&cyc, 8));
        CK(cudaMemcpy(A, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(B, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
        const size_t smem = (size_t)(K / 8) * (128 + 256) * 16;
        mma_probe<<<1, 128, smem>>>(A, B, K, D, Dp, cyc, 1);
        CK(cudaDeviceSynchronize());
        std::vector<float> hD(128 * 256);
        std::vector<uint32_t> hDp(128 * 128);
        CK(cudaMemcpy(hD.data(), D, hD.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hDp.data(), Dp, hDp.size() * 4, cudaMemcpyDeviceToHost));
        double maxdiff = 0, maxdiff_ulp = 0;
        int n_exact = 0, n_floor = 0, n_rn = 0;
        for (int r = 0; r < 128; r++)
            for (int c = 0; c < 256; c++) {
                double ref = 0;
                for (int k = 0; k < K; k++) ref += dA[r * K + k] * dB[c * K + k];  // exact in double (small dyadic values)
                const double got = (double)hD[r * 256 + c];
                const double diff = fabs(got - ref);
                if (diff > maxdiff) maxdiff = diff;
                if (magic > 0) {
                    const double ulp = magic == 49152.0 ? 1.0 / 256 : 1.0 / 32768;
                    maxdiff_ulp = fmax(maxdiff_ulp, diff / ulp);
                    if (got == ref) n_exact++;
                    if (got == floor(ref / ulp) * ulp) n_floor++;
                    if (got == nearbyint(ref / ulp) * ulp) n_rn++;
                }
            }
        printf("MMA test %d (magic %.0f): max |D - exact| = %.6g", test, magic, maxdiff);
        if (magic > 0)
            printf("  = %.3f ulp(magic);  exact %d  ==floor %d  ==round-nearest %d of %d", maxdiff_ulp, n_exact, n_floor, n_rn,
                   128 * 256);
        printf("\n");
        if (test == 1) {
            // pack::16b semantics: compare packed words with the low / high halves of the f32 bit patterns
            int lo_match = 0, hi_match = 0;
            for (int r = 0; r < 128; r++)
                for (int j = 0; j < 128; j++) {
                    uint32_t b0, b1;
                    memcpy(&b0, &hD[r * 256 + 2 * j], 4);
                    memcpy(&b1, &hD[r * 256 + 2 * j + 1], 4);
                    const uint32_t p = hDp[r * 128 + j];
                    if (p == ((b0 & 0xffffu) | (b1 << 16))) lo_match++;
                    if (p == ((b0 >> 16) | (b1 & 0xffff0000u))) hi_match++;
                }
            printf("pack::16b: packed == (lo16(col 2j) | lo16(col 2j+1)<<16) for %d / %d words; == hi16 variant for %d\n",
                   lo_match, 128 * 128, hi_match);
            printf("  sample row 5: f32 bits %08x %08x -> packed %08x\n", *(uint32_t *)&hD[5 * 256], *(uint32_t *)&hD[5 * 256 + 1],
                   hDp[5 * 128]);
        }
        // timing: many back-to-back MMAs
        if (test == 0) {
            for (int reps : {64, 256}) {
                mma_probe<<<1, 128, smem>>>(A, B, K, D, Dp, cyc, reps);
                CK(cudaDeviceSynchronize());
                long long c;
                CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
                printf("MMA timing: %d x (2 x tcgen05.mma M128 N256 K16 f16) = %lld cycles -> %.1f cycles per K=16 instruction\n", reps, c,
                       (double)c / (reps * 2));
            }
        }
        cudaFree(A); cudaFree(B); cudaFree(D); cudaFree(Dp); cudaFree(cyc);
    }

    // ---- 3: tcgen05.ld throughput -----------------------------------------------------------------------
    {
        unsigned *sink;
        long long *cyc_d;
        CK(cudaMalloc(&sink, 148 * 256 * 4));
        CK(cudaMalloc(&cyc_d, 148 * 8));
        const int iters = 2000;
        for (int mode = 0; mode < 3; mode++)
            for (int threads : {128, 256}) {
                cudaEvent_t e0, e1;
                cudaEventCreate(&e0); cudaEventCreate(&e1);
                for (int pass = 0; pass < 2; pass++) {
                    if (pass == 1) cudaEventRecord(e0);
                    if (mode == 0) ldtm_bench<0><<<148, threads>>>(iters, sink, cyc_d);
                    if (mode == 1) ldtm_bench<1><<<148, threads>>>(iters, sink, cyc_d);
                    if (mode == 2) ldtm_bench<2><<<148, threads>>>(iters, sink, cyc_d);
                    if (pass == 1) cudaEventRecord(e1);
                    CK(cudaDeviceSynchronize());
                }
                float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
                std::vector<long long> cyc(148);
                CK(cudaMemcpy(cyc.data(), cyc_d, 148 * 8, cudaMemcpyDeviceToHost));
                double mean = 0;
                for (auto c : cyc) mean += (double)c;
                mean /= 148;
                const double cells = (double)iters * 256 * threads;  // 32-bit TMEM cells read per SM
                printf("LDTM %-12s %d warps: %.1f cells/clk/SM = %.0f B/clk/SM (TMEM bytes)  -> 128x256 tile in %.0f clk64 = %.1f ns (kernel %.3f ms)\n",
                       mode == 0 ? "32x32b.x32" : (mode == 1 ? "x32.pack16" : "32x32b.x16"), threads / 32, cells / mean,
                       cells * 4 / mean, 32768.0 / (cells / mean), 32768.0 / (cells / (ms * 1e6)), ms);
            }
        // ---- 4: ALU rates ---------------------------------------------------------------------------------
        unsigned *sink2;
        long long *cyc2;
        CK(cudaMalloc(&sink2, 148 * 2 * 512 * 4));
        CK(cudaMalloc(&cyc2, 148 * 2 * 8));
        run_alu<FMNMX3>(sink2, cyc2);
        run_alu<FMNMX2>(sink2, cyc2);
        run_alu<VIMNMX3>(sink2, cyc2);
        run_alu<VIMNMX3_16>(sink2, cyc2);
        run_alu<LOP3>(sink2, cyc2);
        run_alu<LEA>(sink2, cyc2);
        run_alu<IMAD>(sink2, cyc2);
        run_alu<FADD>(sink2, cyc2);
        run_alu<FFMA>(sink2, cyc2);
        run_alu<PRMT>(sink2, cyc2);
        run_alu<VIADDMNMX>(sink2, cyc2);
        run_alu<MIX_IMAD_VIMNMX3>(sink2, cyc2);
        run_alu<MIX_FADD_FMNMX3>(sink2, cyc2);
        run_alu<MIX_LOP3_VIMNMX3>(sink2, cyc2);
        run_alu<MIX_LEA_VIMNMX3>(sink2, cyc2);
        run_alu<LEA_IMM>(sink2, cyc2);
        run_alu<LOP3_IMM>(sink2, cyc2);
        run_alu<FADD_IMM>(sink2, cyc2);
        run_alu<MIX_LEAIMM_VIMNMX3>(sink2, cyc2);
        run_alu<MIX_LOP3IMM_VIMNMX3>(sink2, cyc2);
        run_alu<MIX_LOP3IMM_FMNMX3>(sink2, cyc2);
        run_alu<MIX_FADDIMM_FMNMX3>(sink2, cyc2);
    }
    printf("probe done\n");
    return 0;
}
