// How many shared-memory wavefronts (cycles of the LSU data pipe, one per clock per SM) does a 16-byte load cost for a
// given lane -> address pattern?  Each warp issues LDS.128 in a loop; the pattern maps the lane to a 16-byte chunk index.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/lds_probe tools/lds_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ int pattern(int p, int l)
{
    switch (p) {
    case 0: return 0;                                   // one address for the whole warp
    case 1: return l;                                   // 32 distinct contiguous chunks (512 B)
    case 2: return l >> 1;                              // pairs of consecutive lanes share (16 distinct)
    case 3: return l >> 2;                              // groups of 4 consecutive lanes share (8 distinct = 128 B)
    case 4: return l >> 3;                              // groups of 8 consecutive lanes share (4 distinct = 64 B)
    case 5: return l >> 4;                              // half-warps share (2 distinct, adjacent)
    case 6: return l & 7;                               // 8 distinct, period 8
    case 7: return l & 3;                               // 4 distinct, period 4
    case 8: return l & 15;                              // 16 distinct, period 16 (config 3 A fragment, linear mapping)
    case 9: return (l & 3) + 4 * (l >> 4);              // 4 distinct per half-warp, period 4 (A fragment, patch mapping)
    case 10: return ((l >> 2) & 3) * 64 + ((l >> 2) & 3);   // B fragment, patch mapping: 4 chunks 1 KiB apart, swizzled; both half-warps alike
    case 11: return (l >> 4) * 64;                      // B fragment, linear mapping: half-warps 1 KiB apart (same bank)
    case 12: return (l >> 2) & 3;                       // groups of 4 share, 4 distinct, both half-warps alike
    case 13: return (l >> 1) & 7;                       // pairs share, 8 distinct (128 B), both half-warps alike
    case 14: return (l & 1) + 2 * (l >> 3);             // 2 distinct per quarter-warp, period 2
    case 15: return (l & 7) ^ (l >> 3);                 // 8 distinct per quarter, permuted per quarter (same 128 B)
    case 16: return (l & 7) + 8 * (l >> 4);             // quarter pairs alike: Q0 = Q1 (128 B), Q2 = Q3 (next 128 B)
    case 17: return (l & 7) + 8 * ((l >> 3) & 1);       // half-warps alike, 16 distinct per half (256 B)
    default: return 0;
    }
}

__global__ void __launch_bounds__(256) k_lds(uint32_t *out, int p, int iters)
{
    __shared__ int4 buf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_int4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int4 *ptr = buf + pattern(p, lane) + (threadIdx.x >> 5) * 8;
    int4 acc = make_int4(0, 0, 0, 0);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            int4 v;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"((uint32_t)__cvta_generic_to_shared(ptr) + ((u * 131 + i) & 7) * 1024));
            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
        }
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x7fffffff) out[0] = 1;
}

int main()
{
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    uint32_t *out; cudaMalloc(&out, 64);
    const int iters = 4096, grid = sms * 4;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int p = 0; p <= 17; ++p) {
        k_lds<<<grid, 256>>>(out, p, 64);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        k_lds<<<grid, 256>>>(out, p, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double warp_instr_per_sm = 4.0 * 8 * iters * 8;       // CTAs per SM x warps x iterations x loads
        const double cycles = ms * 1e-3 * khz * 1e3;
        printf("{\"pattern\": %d, \"ms\": %.4f, \"cycles_per_lds128_per_sm\": %.3f}\n", p, ms, cycles / warp_instr_per_sm);
    }
    return cudaDeviceSynchronize() != cudaSuccess;
}
