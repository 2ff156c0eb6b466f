// Issue-rate probe for the pipes the config-3 kernel uses (sm_100a): one instruction kind (or a mix) in 8 independent
// chains per thread, 512 threads x 2 CTAs per SM.  Prints thread-instructions per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/pipe_probe tools/pipe_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITERS 4096

template <int KIND> __device__ __forceinline__ void step(uint32_t (&r)[CHAINS], uint64_t (&q)[CHAINS], uint32_t x, uint32_t y)
{
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
        if constexpr (KIND == 0) asm volatile("fma.rn.f32 %0, %0, %1, 0f4FC00000;" : "+r"(r[c]) : "r"(x));            // FFMA imm
        if constexpr (KIND == 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(r[c]) : "r"(x), "r"(y));             // FFMA reg
        if constexpr (KIND == 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(q[c]) : "l"(q[(c + 1) % CHAINS] ), "l"(q[(c + 2) % CHAINS]));   // FFMA2 (dependent mix; see KIND 12)
        if constexpr (KIND == 3) asm volatile("add.rn.f32 %0, %0, %1;" : "+r"(r[c]) : "r"(x));                         // FADD
        if constexpr (KIND == 4) asm volatile("add.s32 %0, %0, %1;" : "+r"(r[c]) : "r"(x));                            // IADD3 (2 inputs)
        if constexpr (KIND == 5) asm volatile("{.reg .s32 t; add.s32 t, %0, %1; add.s32 %0, t, %2;}" : "+r"(r[c]) : "r"(x), "r"(y));   // -> IADD3 3-input
        if constexpr (KIND == 6) asm volatile("shr.s32 %0, %0, 1;" : "+r"(r[c]));                                       // SHF
        if constexpr (KIND == 7) asm volatile("{.reg .s32 t; shr.s32 t, %0, 1; add.s32 %0, t, %1;}" : "+r"(r[c]) : "r"(x));            // LEA.HI.SX32
        if constexpr (KIND == 8) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(r[c]) : "r"(x), "r"(y));            // IMAD
        if constexpr (KIND == 9) asm volatile("mad.hi.s32 %0, %0, %1, %2;" : "+r"(r[c]) : "r"(x), "r"(y));            // IMAD.HI
        if constexpr (KIND == 10) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(q[c]) : "l"(q[(c + 1) % CHAINS]));    // FADD2
        if constexpr (KIND == 11) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[c]) : "r"(x), "r"(y));       // LOP3
    }
}

template <int KIND> __global__ void __launch_bounds__(512, 2) k_probe(uint32_t *out, uint32_t x, uint32_t y, int iters)
{
    uint32_t r[CHAINS]; uint64_t q[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { r[c] = threadIdx.x + c; q[c] = ((uint64_t)(threadIdx.x + c) << 32) | c; }
    for (int i = 0; i < iters; ++i) step<KIND>(r, q, x, y);
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += r[c] + (uint32_t)q[c] + (uint32_t)(q[c] >> 32);
    if (s == 0x12345) out[0] = s;
}

// mixes: NA instructions of kind A then NB of kind B per chain step
template <int KA, int NA, int KB, int NB> __global__ void __launch_bounds__(512, 2) k_mix(uint32_t *out, uint32_t x, uint32_t y, int iters)
{
    uint32_t r[CHAINS], r2[CHAINS]; uint64_t q[CHAINS], q2[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { r[c] = threadIdx.x + c; r2[c] = r[c] * 3; q[c] = ((uint64_t)(threadIdx.x + c) << 32) | c; q2[c] = q[c] + 5; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int a = 0; a < NA; ++a) step<KA>(r, q, x, y);
#pragma unroll
        for (int b = 0; b < NB; ++b) step<KB>(r2, q2, x, y);
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += r[c] + r2[c] + (uint32_t)q[c] + (uint32_t)(q[c] >> 32) + (uint32_t)q2[c] + (uint32_t)(q2[c] >> 32);
    if (s == 0x12345) out[0] = s;
}

template <class F> static void run(const char *name, F launch, double instr_per_iter_per_thread, int sms, double mhz)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(256);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch(ITERS);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double threads = (double)sms * 2 * 512;
    const double instr = threads * instr_per_iter_per_thread * ITERS;
    printf("{\"probe\": \"%s\", \"ms\": %.4f, \"T_thread_instr_per_s\": %.3f, \"thread_instr_per_clk_per_sm\": %.2f}\n", name, ms, instr / ms / 1e9, instr / (ms * 1e-3) / (mhz * 1e6) / sms);
}

int main()
{
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    uint32_t *out; cudaMalloc(&out, 64);
    printf("{\"sms\": %d, \"clock_mhz\": %.0f}\n", sms, mhz);
    const int grid = sms * 2;
#define PROBE(K, NAME) run(NAME, [&](int it) { k_probe<K><<<grid, 512>>>(out, 3u, 5u, it); }, CHAINS, sms, mhz)
#define MIX(KA, NA, KB, NB, NAME) run(NAME, [&](int it) { k_mix<KA, NA, KB, NB><<<grid, 512>>>(out, 3u, 5u, it); }, CHAINS * (NA + NB), sms, mhz)
    PROBE(0, "FFMA imm"); PROBE(1, "FFMA reg"); PROBE(2, "FFMA2 (f32x2, counts 1 instr)"); PROBE(3, "FADD"); PROBE(10, "FADD2 (f32x2)");
    PROBE(4, "IADD 2-input"); PROBE(5, "IADD3 3-input"); PROBE(6, "SHF"); PROBE(7, "LEA.HI (shr+add)"); PROBE(8, "IMAD"); PROBE(9, "IMAD.HI"); PROBE(11, "LOP3");
    MIX(0, 1, 5, 1, "FFMAimm:IADD3 1:1"); MIX(0, 2, 5, 1, "FFMAimm:IADD3 2:1"); MIX(0, 1, 7, 1, "FFMAimm:LEA.HI 1:1");
    MIX(2, 1, 5, 1, "FFMA2:IADD3 1:1"); MIX(2, 1, 5, 2, "FFMA2:IADD3 1:2"); MIX(8, 1, 5, 1, "IMAD:IADD3 1:1"); MIX(1, 1, 5, 1, "FFMAreg:IADD3 1:1");
    MIX(3, 1, 5, 1, "FADD:IADD3 1:1"); MIX(0, 1, 3, 1, "FFMAimm:FADD 1:1"); MIX(5, 1, 7, 1, "IADD3:LEA.HI 1:1");
    cudaError_t e = cudaDeviceSynchronize();
    printf("{\"status\": \"%s\"}\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
