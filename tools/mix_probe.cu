// Upper bound of the config-3 inner loop from registers only (no shared memory, no barriers): per group of four MACs
// 4 FFMA (immediate addend), IADD3 (three registers), IADD3 (two registers + immediate), SHF, LEA.HI -- the instruction
// mix of the deferred-shift tree -- in CH independent chains per thread, two instruction orders.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/mix_probe tools/mix_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define FFMA(d, a, b) asm volatile("fma.rn.f32 %0, %1, %2, 0f4F400000;" : "=r"(d) : "r"(a), "r"(b))
#define ADD3(d, a, b, c) asm volatile("{.reg .s32 t; add.s32 t, %1, %2; add.s32 %0, t, %3;}" : "=r"(d) : "r"(a), "r"(b), "r"(c))
#define ADD2I(d, a, b) asm volatile("{.reg .s32 t; add.s32 t, %1, %2; add.s32 %0, t, -1023410177;}" : "=r"(d) : "r"(a), "r"(b))
#define SHR1(d, a) asm volatile("shr.s32 %0, %1, 1;" : "=r"(d) : "r"(a))
#define SHRADD(d, a, b) asm volatile("{.reg .s32 t; shr.s32 t, %1, 1; add.s32 %0, t, %2;}" : "=r"(d) : "r"(a), "r"(b))

template <int CH, int ORDER> __global__ void __launch_bounds__(256, 2) k_mix(uint32_t *out, const uint32_t *in, int iters)
{
    uint32_t a[4], a0[4], b[CH], acc[CH], half[CH];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = a0[i] = in[threadIdx.x + i * 256];
#pragma unroll
    for (int c = 0; c < CH; ++c) { b[c] = in[1024 + threadIdx.x + c * 256]; acc[c] = 0; half[c] = 0; }
    for (int it = 0; it < iters; ++it) {
        uint32_t p[CH][4], s[CH], t[CH];
        if (ORDER == 0) {                                   // by kind: all products, all sums, all shifts
#pragma unroll
            for (int c = 0; c < CH; ++c)
#pragma unroll
                for (int k = 0; k < 4; ++k) FFMA(p[c][k], a[k], b[c]);
#pragma unroll
            for (int c = 0; c < CH; ++c) ADD3(s[c], p[c][0], p[c][1], p[c][2]);
#pragma unroll
            for (int c = 0; c < CH; ++c) ADD2I(s[c], s[c], p[c][3]);
#pragma unroll
            for (int c = 0; c < CH; ++c) SHR1(t[c], half[c]);
#pragma unroll
            for (int c = 0; c < CH; ++c) { SHRADD(acc[c], s[c], t[c]); half[c] = acc[c]; }
        } else {                                            // interleaved: chain c's adds behind chain c+1's products
#pragma unroll
            for (int c = 0; c < CH; ++c) {
#pragma unroll
                for (int k = 0; k < 4; ++k) FFMA(p[c][k], a[k], b[c]);
                if (c > 0) { ADD3(s[c - 1], p[c - 1][0], p[c - 1][1], p[c - 1][2]); SHR1(t[c - 1], half[c - 1]); }
                if (c > 1) { ADD2I(s[c - 2], s[c - 2], p[c - 2][3]); SHRADD(acc[c - 2], s[c - 2], t[c - 2]); half[c - 2] = acc[c - 2]; }
            }
            ADD3(s[CH - 1], p[CH - 1][0], p[CH - 1][1], p[CH - 1][2]); SHR1(t[CH - 1], half[CH - 1]);
            ADD2I(s[CH - 2], s[CH - 2], p[CH - 2][3]); SHRADD(acc[CH - 2], s[CH - 2], t[CH - 2]); half[CH - 2] = acc[CH - 2];
            ADD2I(s[CH - 1], s[CH - 1], p[CH - 1][3]); SHRADD(acc[CH - 1], s[CH - 1], t[CH - 1]); half[CH - 1] = acc[CH - 1];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) a[k] = a0[k] ^ ((uint32_t)it << (8 + k));   // loop-variant operands (4 extra ALU ops per 4 * CH MACs)
    }
    uint32_t r = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) r ^= acc[c];
    out[blockIdx.x * 256 + threadIdx.x] = r;
}

// Variant C: the second add of a group on the FMA pipe.  IMAD (d = s * 1 + p3, the 1 in a register the compiler cannot see
// through) issues on the heavy half of the FMA pipe next to the ALU pipe (tools/pipe_probe.cu: IMAD:IADD3 1:1 = 115 per clock),
// the group constant moves into the magic constant of one FFMA, so the instruction count stays what it is: per four MACs
// 4 FFMA + IADD3 (three registers) + IMAD + SHF + LEA.HI.  IMADS > 1: further adds moved (IMADS = 2: the merge's add).
#define IMAD1(d, a, one, c) asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(c))
#define FFMAB(d, a, b) asm volatile("fma.rn.f32 %0, %1, %2, 0f4F3FFFC0;" : "=r"(d) : "r"(a), "r"(b))
template <int CH, int ORDER> __global__ void __launch_bounds__(256, 2) k_mix_imad(uint32_t *out, const uint32_t *in, int iters)
{
    uint32_t a[4], a0[4], b[CH], acc[CH], half[CH];
    const uint32_t one = in[2048 + (threadIdx.x & 1)];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = a0[i] = in[threadIdx.x + i * 256];
#pragma unroll
    for (int c = 0; c < CH; ++c) { b[c] = in[1024 + threadIdx.x + c * 256]; acc[c] = 0; half[c] = 0; }
    for (int it = 0; it < iters; ++it) {
        uint32_t p[CH][4], s[CH], t[CH];
        if (ORDER == 0) {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
#pragma unroll
                for (int k = 0; k < 3; ++k) FFMA(p[c][k], a[k], b[c]);
                FFMAB(p[c][3], a[3], b[c]);
            }
#pragma unroll
            for (int c = 0; c < CH; ++c) ADD3(s[c], p[c][0], p[c][1], p[c][2]);
#pragma unroll
            for (int c = 0; c < CH; ++c) IMAD1(s[c], s[c], one, p[c][3]);
#pragma unroll
            for (int c = 0; c < CH; ++c) SHR1(t[c], half[c]);
#pragma unroll
            for (int c = 0; c < CH; ++c) { SHRADD(acc[c], s[c], t[c]); half[c] = acc[c]; }
        } else {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
#pragma unroll
                for (int k = 0; k < 3; ++k) FFMA(p[c][k], a[k], b[c]);
                FFMAB(p[c][3], a[3], b[c]);
                if (c > 0) { ADD3(s[c - 1], p[c - 1][0], p[c - 1][1], p[c - 1][2]); SHR1(t[c - 1], half[c - 1]); }
                if (c > 1) { IMAD1(s[c - 2], s[c - 2], one, p[c - 2][3]); SHRADD(acc[c - 2], s[c - 2], t[c - 2]); half[c - 2] = acc[c - 2]; }
            }
            ADD3(s[CH - 1], p[CH - 1][0], p[CH - 1][1], p[CH - 1][2]); SHR1(t[CH - 1], half[CH - 1]);
            IMAD1(s[CH - 2], s[CH - 2], one, p[CH - 2][3]); SHRADD(acc[CH - 2], s[CH - 2], t[CH - 2]); half[CH - 2] = acc[CH - 2];
            IMAD1(s[CH - 1], s[CH - 1], one, p[CH - 1][3]); SHRADD(acc[CH - 1], s[CH - 1], t[CH - 1]); half[CH - 1] = acc[CH - 1];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) a[k] = a0[k] ^ ((uint32_t)it << (8 + k));
    }
    uint32_t r = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) r ^= acc[c];
    out[blockIdx.x * 256 + threadIdx.x] = r;
}
// Variant D: the merge as two shifts (ALU) and an IMAD add; the group as in variant C: 4 FFMA + IADD3 + 2 IMAD + 2 SHF
template <int CH> __global__ void __launch_bounds__(256, 2) k_mix_imad2(uint32_t *out, const uint32_t *in, int iters)
{
    uint32_t a[4], a0[4], b[CH], acc[CH], half[CH];
    const uint32_t one = in[2048 + (threadIdx.x & 1)];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = a0[i] = in[threadIdx.x + i * 256];
#pragma unroll
    for (int c = 0; c < CH; ++c) { b[c] = in[1024 + threadIdx.x + c * 256]; acc[c] = 0; half[c] = 0; }
    for (int it = 0; it < iters; ++it) {
        uint32_t p[CH][4], s[CH], t[CH], u[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
#pragma unroll
            for (int k = 0; k < 3; ++k) FFMA(p[c][k], a[k], b[c]);
            FFMAB(p[c][3], a[3], b[c]);
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) ADD3(s[c], p[c][0], p[c][1], p[c][2]);
#pragma unroll
        for (int c = 0; c < CH; ++c) IMAD1(s[c], s[c], one, p[c][3]);
#pragma unroll
        for (int c = 0; c < CH; ++c) { SHR1(t[c], half[c]); SHR1(u[c], s[c]); }
#pragma unroll
        for (int c = 0; c < CH; ++c) { IMAD1(acc[c], u[c], one, t[c]); half[c] = acc[c]; }
#pragma unroll
        for (int k = 0; k < 4; ++k) a[k] = a0[k] ^ ((uint32_t)it << (8 + k));
    }
    uint32_t r = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) r ^= acc[c];
    out[blockIdx.x * 256 + threadIdx.x] = r;
}

// Variant B: group sums on the FP pipe.  Products with alternating +-C magic constants make F0 + F1 exact, so a group of
// four is 4 FFMA + 3 FADD and ONE float->int conversion with the level-1 rounding built in (cvt.rpi: ceil(s / 2)); the
// level-2 node is one IADD3 (two registers + immediate) per two groups, the levels above cost SHF + LEA.HI per merge
// (7 merges per 16 groups).
#define FFMAN(d, a, b) asm volatile("fma.rn.f32 %0, %1, %2, 0fCF400000;" : "=r"(d) : "r"(a), "r"(b))
#define FADD(d, a, b) asm volatile("add.rn.f32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b))
#define F2ICEIL(d, a) asm volatile("cvt.rpi.s32.f32 %0, %1;" : "=r"(d) : "r"(a))
template <int CH> __global__ void __launch_bounds__(256, 2) k_mix_fp(uint32_t *out, const uint32_t *in, int iters)
{
    uint32_t a[4], a0[4], b[CH], acc[CH], half[CH];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = a0[i] = in[threadIdx.x + i * 256];
#pragma unroll
    for (int c = 0; c < CH; ++c) { b[c] = in[1024 + threadIdx.x + c * 256]; acc[c] = 0; half[c] = 0; }
    for (int it = 0; it < iters; it += 2) {                 // two groups per chain and trip: one level-2 node
        uint32_t v[2][CH];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            uint32_t p[CH][4], t01[CH], t23[CH], t[CH];
#pragma unroll
            for (int c = 0; c < CH; ++c) { FFMA(p[c][0], a[0], b[c]); FFMAN(p[c][1], a[1], b[c]); FFMA(p[c][2], a[2], b[c]); FFMAN(p[c][3], a[3], b[c]); }
#pragma unroll
            for (int c = 0; c < CH; ++c) { FADD(t01[c], p[c][0], p[c][1]); FADD(t23[c], p[c][2], p[c][3]); }
#pragma unroll
            for (int c = 0; c < CH; ++c) { FADD(t[c], t01[c], t23[c]); F2ICEIL(v[g][c], t[c]); }
#pragma unroll
            for (int k = 0; k < 4; ++k) a[k] = a0[k] ^ ((uint32_t)(it + g) << (8 + k));
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            uint32_t s2, tt;
            ADD2I(s2, v[0][c], v[1][c]);                    // level-2 pre-shift sum with its bias
            SHR1(tt, half[c]); SHRADD(acc[c], s2, tt);      // one merge above (7 per 8 level-2 nodes in the kernel)
            half[c] = acc[c];
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) r ^= acc[c];
    out[blockIdx.x * 256 + threadIdx.x] = r;
}

int main()
{
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    uint32_t *out, *in; cudaMalloc(&out, 4 << 20); cudaMalloc(&in, 1 << 20); cudaMemset(in, 0x3f, 1 << 20); { const uint32_t ones[2] = { 1u, 1u }; cudaMemcpy(in + 2048, ones, sizeof ones, cudaMemcpyHostToDevice); }
    const int iters = 8192, grid = sms * 2;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char *name, auto kern, int ch) {
        kern<<<grid, 256>>>(out, in, 64);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        kern<<<grid, 256>>>(out, in, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) printf("{\"error\": \"%s\"}\n", cudaGetErrorString(err));
        const double macs = (double)grid * 256 * iters * ch * 4;
        printf("{\"probe\": \"%s\", \"ms\": %.4f, \"mac_per_clk_per_sm\": %.2f, \"tmacs_equiv\": %.2f}\n", name, ms, macs / (ms * 1e-3) / (khz * 1e3) / sms, macs / ms / 1e9);
    };
    run("8 chains, by kind", k_mix<8, 0>, 8); run("8 chains, interleaved", k_mix<8, 1>, 8);
    run("16 chains, by kind", k_mix<16, 0>, 16); run("16 chains, interleaved", k_mix<16, 1>, 16);
    run("IMAD second add, 8 chains, by kind", k_mix_imad<8, 0>, 8); run("IMAD second add, 8 chains, interleaved", k_mix_imad<8, 1>, 8);
    run("IMAD second add, 16 chains, by kind", k_mix_imad<16, 0>, 16); run("IMAD second add, 16 chains, interleaved", k_mix_imad<16, 1>, 16);
    run("IMAD second add + IMAD merge add, 8 chains", k_mix_imad2<8>, 8); run("IMAD second add + IMAD merge add, 16 chains", k_mix_imad2<16>, 16);
    run("FP group sums + F2I, 8 chains", k_mix_fp<8>, 8); run("FP group sums + F2I, 16 chains", k_mix_fp<16>, 16);
    return cudaDeviceSynchronize() != cudaSuccess;
}
