// qb_fast_reduce.cuh — compile-time-specialised Qreduce (vector form, QuBLAS.h:4960-4990) for
// power-of-two vectors (BASELINE config 2: 2^24 x Qu<8,8>, per-level TypeList).  HBM-bound:
// 4 algorithmic bytes per element, every level's quantizer is a couple of immediates.
//
//   lane  : 8 contiguous words (two 128-bit loads)          -> levels LEV0 .. LEV0+2 in registers
//   warp  : 256 contiguous words                             -> levels +3 .. +7 by xor-shuffle (adjacent pairs)
//   chunk : 2^(PW-8) such groups, 4 groups of loads in flight -> levels +8 .. via the level-indexed register stack
// Passes are chained (tree separability): pass 1 leaves one level-PW node per chunk,
// pass 2 starts at LEV0 = PW on those nodes.
#pragma once
#include "qb_device.cuh"

namespace qb {

template <class Cfg, int BASE, int L, int NTOP, typename I>
QB_DEV void rstack_push(I (&st)[NTOP > 0 ? NTOP : 1], uint32_t c, I v, I &result)
{
    if constexpr (L >= NTOP) {
        result = v;
    } else {
        if ((c >> L) & 1) {
            constexpr AddStage lv = Cfg::lvl(BASE + L);
            I m = merge_apply<I, true>(st[L], v, lv);
            rstack_push<Cfg, BASE, L + 1, NTOP, I>(st, c, m, result);
        } else {
            st[L] = v;
        }
    }
}

template <typename I> QB_DEV I shfl_xor_ct(I v, int m)
{
    if constexpr (sizeof(I) == 8) return (I)__shfl_xor_sync(0xffffffffu, (long long)v, m);
    else return (I)__shfl_xor_sync(0xffffffffu, (int)v, m);
}

// x: batch * n_in words (int32), n_in a multiple of 2^PW.  out: one node per chunk (int32).
template <class Cfg, int LEV0, int PW>
QB_DEV void reduce_ct_body(const int32_t *__restrict__ x, int32_t *__restrict__ out, uint64_t total_chunks)
{
    typedef typename Cfg::I I;
    constexpr int NSUB = 1 << (PW - 8), NTOP = PW - 8, U = NSUB < 4 ? NSUB : 4;
    const int lane = threadIdx.x & 31;
    const uint64_t warps_in_grid = (uint64_t)gridDim.x * (blockDim.x >> 5);
    constexpr AddStage l0 = Cfg::lvl(LEV0), l1 = Cfg::lvl(LEV0 + 1), l2 = Cfg::lvl(LEV0 + 2);
    for (uint64_t chunk = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); chunk < total_chunks; chunk += warps_in_grid) {
        const int4 *src = (const int4 *)(x + (chunk << PW)) + lane * 2;
        I st[NTOP > 0 ? NTOP : 1];
        I result = 0;
#pragma unroll 1
        for (int j0 = 0; j0 < NSUB; j0 += U) {
            int4 q[U][2];
#pragma unroll
            for (int u = 0; u < U; u++) {
                q[u][0] = __ldcs(src + (size_t)(j0 + u) * 64);          // streaming loads: every word is read exactly once
                q[u][1] = __ldcs(src + (size_t)(j0 + u) * 64 + 1);
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                I a0 = merge_apply<I, true>((I)q[u][0].x, (I)q[u][0].y, l0);
                I a1 = merge_apply<I, true>((I)q[u][0].z, (I)q[u][0].w, l0);
                I a2 = merge_apply<I, true>((I)q[u][1].x, (I)q[u][1].y, l0);
                I a3 = merge_apply<I, true>((I)q[u][1].z, (I)q[u][1].w, l0);
                I b0 = merge_apply<I, true>(a0, a1, l1);
                I b1 = merge_apply<I, true>(a2, a3, l1);
                I cur = merge_apply<I, true>(b0, b1, l2);
                { constexpr AddStage s = Cfg::lvl(LEV0 + 3); cur = merge_apply<I, true>(cur, shfl_xor_ct<I>(cur, 1), s); }
                { constexpr AddStage s = Cfg::lvl(LEV0 + 4); cur = merge_apply<I, true>(cur, shfl_xor_ct<I>(cur, 2), s); }
                { constexpr AddStage s = Cfg::lvl(LEV0 + 5); cur = merge_apply<I, true>(cur, shfl_xor_ct<I>(cur, 4), s); }
                { constexpr AddStage s = Cfg::lvl(LEV0 + 6); cur = merge_apply<I, true>(cur, shfl_xor_ct<I>(cur, 8), s); }
                { constexpr AddStage s = Cfg::lvl(LEV0 + 7); cur = merge_apply<I, true>(cur, shfl_xor_ct<I>(cur, 16), s); }
                rstack_push<Cfg, LEV0 + 8, 0, NTOP, I>(st, (uint32_t)(j0 + u), cur, result);
            }
        }
        if (lane == 0) out[chunk] = (int32_t)result;
    }
}

template <class Cfg, int LEV0, int PW>
__global__ void __launch_bounds__(256)
k_reduce_ct(const int32_t *__restrict__ x, int32_t *__restrict__ out, uint64_t total_chunks)
{
    reduce_ct_body<Cfg, LEV0, PW>(x, out, total_chunks);
}

#ifndef QB_JIT
// Two chained passes for len = 2^LOGN (LOGN in [16, 24]): pass 1 PW1 = 12, pass 2 covers the rest.
template <class Cfg>
static int launch_reduce_ct(const void *x, void *out, void *scratch, uint64_t batch, cudaStream_t s)
{
    constexpr int LOGN = Cfg::LOGN, PW1 = 12, PW2 = LOGN - PW1;
    static_assert(PW2 >= 8 && PW2 <= 12, "supported lengths: 2^20 .. 2^24");
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    const uint64_t chunks1 = batch << PW2;
    const uint64_t blocks1 = (chunks1 + 7) / 8;
    const unsigned grid1 = (unsigned)(blocks1 < (uint64_t)sms * 8 ? blocks1 : (uint64_t)sms * 8);
    k_reduce_ct<Cfg, 0, PW1><<<grid1, 256, 0, s>>>((const int32_t *)x, (int32_t *)scratch, chunks1);
    const uint64_t blocks2 = (batch + 7) / 8;
    k_reduce_ct<Cfg, PW1, PW2><<<(unsigned)(blocks2 < 1184 ? blocks2 : 1184), 256, 0, s>>>((const int32_t *)scratch, (int32_t *)out, batch);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) count_launch(2);
    return (int)e;
}

#endif  // QB_JIT

}  // namespace qb
