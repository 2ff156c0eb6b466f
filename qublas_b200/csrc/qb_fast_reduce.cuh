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
        // fully unrolled (<= 4 passes): the stack branches on the group counter fold and all loads of a chunk can be in
        // flight at once; measured on config 2 (NVRTC build): unroll 1 / 2 / 4 = 1 499 / 1 539 / 1 663 Gelem/s
#ifndef QB_REDUCE_UNROLL
#define QB_REDUCE_UNROLL 0
#endif
        constexpr int GROUP_UNROLL = QB_REDUCE_UNROLL > 0 ? QB_REDUCE_UNROLL : (NSUB / U > 0 ? NSUB / U : 1);
#pragma unroll GROUP_UNROLL
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

// ---------------------------------------------------------------------------------------------------------------
// Any length (vector form): the same lane / warp / chunk decomposition, one warp per chunk of 2^PW leaves.  Chunks
// that are complete take the path above; the last chunk of a vector whose length is not a multiple of 2^PW tracks
// which nodes exist and, where a right-hand partner is missing, converts the leftover into the level's type exactly
// as the reference carries it upward (QuBLAS.h:4977-4980).  RMAX = number of levels this pass may apply (PW, or
// fewer in the final pass, where the tree ends at ceil(log2 n)).
// ---------------------------------------------------------------------------------------------------------------
template <class Cfg, int LEV0, int RMAX, int R, typename I>
QB_DEV void comb_ct(I &lv, bool &lp, I rv, bool rp)
{
    if constexpr (R < RMAX) {
        if (!lp) return;
        if (rp) { constexpr AddStage s = Cfg::lvl(LEV0 + R); lv = merge_apply<I, true>(lv, rv, s); }
        else    { constexpr QStage c = Cfg::carry(LEV0 + R); lv = conv_apply<I, true>(lv, c); }
    }
}

template <class Cfg, int LEV0, int RMAX, int S, typename I>
QB_DEV void comb_shfl_ct(I &cur, bool &cp, int lane)
{
    const I ov = shfl_xor_ct<I>(cur, 1 << S);
    const bool op = __shfl_xor_sync(0xffffffffu, (int)cp, 1 << S) != 0;
    const bool upper = (lane >> S) & 1;
    I lv = upper ? ov : cur, rv = upper ? cur : ov;
    bool lp = upper ? op : cp, rp = upper ? cp : op;
    comb_ct<Cfg, LEV0, RMAX, 3 + S, I>(lv, lp, rv, rp);
    cur = lv; cp = lp;
}

template <class Cfg, int LEV0, int RMAX, int L, int NTOP, typename I>
QB_DEV void pstack_push(I (&st)[NTOP + 1], bool (&sp)[NTOP + 1], uint32_t j, I cur, bool cp)
{
    if constexpr (L >= NTOP) {
        st[L] = cur; sp[L] = cp;
    } else {
        if ((j >> L) & 1) {
            I lv = st[L]; bool lp = sp[L];
            comb_ct<Cfg, LEV0, RMAX, 8 + L, I>(lv, lp, cur, cp);
            pstack_push<Cfg, LEV0, RMAX, L + 1, NTOP, I>(st, sp, j, lv, lp);
        } else {
            st[L] = cur; sp[L] = cp;
        }
    }
}

// x: batch vectors of n_in words, in_stride words apart; out: ceil(n_in / 2^PW) nodes per vector, out_stride words apart.
// (The variadic form of a non-power-of-two length hands over only the whole chunks of each vector: n_in < in_stride.)
template <class Cfg, int LEV0, int PW, int RMAX>
QB_DEV void reduce_any_body(const int32_t *__restrict__ x, int32_t *__restrict__ out, uint64_t n_in, uint64_t batch, int vec_ok,
                            uint64_t in_stride, uint64_t out_stride)
{
    typedef typename Cfg::I I;
    constexpr int NSUB = 1 << (PW - 8), NTOP = PW - 8, U = NSUB < 4 ? NSUB : 4;
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t per_vec = (n_in + ((uint64_t)1 << PW) - 1) >> PW;
    if (warp >= batch * per_vec) return;
    const uint64_t b = warp / per_vec, c = warp - b * per_vec;
    const int32_t *xv = x + b * in_stride;
    const uint64_t e_begin = c << PW;
    const uint64_t e_end = (e_begin + ((uint64_t)1 << PW)) < n_in ? e_begin + ((uint64_t)1 << PW) : n_in;
    I result = 0;
    if (RMAX >= PW && e_end - e_begin == ((uint64_t)1 << PW)) {
        if constexpr (RMAX >= PW) {
            constexpr AddStage l0 = Cfg::lvl(LEV0), l1 = Cfg::lvl(LEV0 + 1), l2 = Cfg::lvl(LEV0 + 2);
            I st[NTOP > 0 ? NTOP : 1];
            constexpr int GROUP_UNROLL = NSUB / U > 0 ? NSUB / U : 1;          // as in reduce_ct_body
#pragma unroll GROUP_UNROLL
            for (int j0 = 0; j0 < NSUB; j0 += U) {
                I v[U][8];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int32_t *p = xv + e_begin + (uint64_t)(j0 + u) * 256 + lane * 8;
                    if (vec_ok) {
                        const int4 q0 = __ldcs((const int4 *)p), q1 = __ldcs((const int4 *)p + 1);
                        v[u][0] = q0.x; v[u][1] = q0.y; v[u][2] = q0.z; v[u][3] = q0.w; v[u][4] = q1.x; v[u][5] = q1.y; v[u][6] = q1.z; v[u][7] = q1.w;
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; i++) v[u][i] = __ldcs(p + i);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    I a0 = merge_apply<I, true>(v[u][0], v[u][1], l0), a1 = merge_apply<I, true>(v[u][2], v[u][3], l0);
                    I a2 = merge_apply<I, true>(v[u][4], v[u][5], l0), a3 = merge_apply<I, true>(v[u][6], v[u][7], l0);
                    I cur = merge_apply<I, true>(merge_apply<I, true>(a0, a1, l1), merge_apply<I, true>(a2, a3, l1), l2);
                    { constexpr AddStage s = Cfg::lvl(LEV0 + 3); cur = merge_apply<I, true>(cur, shfl_xor_ct<I>(cur, 1), s); }
                    { constexpr AddStage s = Cfg::lvl(LEV0 + 4); cur = merge_apply<I, true>(cur, shfl_xor_ct<I>(cur, 2), s); }
                    { constexpr AddStage s = Cfg::lvl(LEV0 + 5); cur = merge_apply<I, true>(cur, shfl_xor_ct<I>(cur, 4), s); }
                    { constexpr AddStage s = Cfg::lvl(LEV0 + 6); cur = merge_apply<I, true>(cur, shfl_xor_ct<I>(cur, 8), s); }
                    { constexpr AddStage s = Cfg::lvl(LEV0 + 7); cur = merge_apply<I, true>(cur, shfl_xor_ct<I>(cur, 16), s); }
                    rstack_push<Cfg, LEV0 + 8, 0, NTOP, I>(st, (uint32_t)(j0 + u), cur, result);
                }
            }
        }
    } else {
        I st[NTOP + 1]; bool sp[NTOP + 1];
#pragma unroll 1
        for (uint32_t j = 0; j < (uint32_t)NSUB; j++) {
            const uint64_t e0 = e_begin + (uint64_t)j * 256 + lane * 8;
            I v[8]; bool p[8];
#pragma unroll
            for (int i = 0; i < 8; i++) { p[i] = e0 + i < e_end; v[i] = p[i] ? (I)__ldcs(xv + e0 + i) : (I)0; }
#pragma unroll
            for (int i = 0; i < 4; i++) comb_ct<Cfg, LEV0, RMAX, 0, I>(v[2 * i], p[2 * i], v[2 * i + 1], p[2 * i + 1]);
            comb_ct<Cfg, LEV0, RMAX, 1, I>(v[0], p[0], v[2], p[2]);
            comb_ct<Cfg, LEV0, RMAX, 1, I>(v[4], p[4], v[6], p[6]);
            comb_ct<Cfg, LEV0, RMAX, 2, I>(v[0], p[0], v[4], p[4]);
            I cur = v[0]; bool cp = p[0];
            comb_shfl_ct<Cfg, LEV0, RMAX, 0, I>(cur, cp, lane);
            comb_shfl_ct<Cfg, LEV0, RMAX, 1, I>(cur, cp, lane);
            comb_shfl_ct<Cfg, LEV0, RMAX, 2, I>(cur, cp, lane);
            comb_shfl_ct<Cfg, LEV0, RMAX, 3, I>(cur, cp, lane);
            comb_shfl_ct<Cfg, LEV0, RMAX, 4, I>(cur, cp, lane);
            pstack_push<Cfg, LEV0, RMAX, 0, NTOP, I>(st, sp, j, cur, cp);
        }
        result = st[NTOP];
    }
    if (lane == 0) out[b * out_stride + c] = (int32_t)result;
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
