// qb_fast_gemul_thread.cuh — compile-time-specialised Qgemul for shapes the register-tile kernel does not take
// (K not a power of two, M or N not a multiple of 4, fewer than 32 tile threads, transposed operands, 1/2/8-byte words,
// the 64-bit register class, both tree forms).  One thread per output element, like the runtime-descriptor kernel
// k_gemul_thread, but every quantizer, the level table and the flush program of the tree (QuBLAS.h:4936-4949 variadic,
// :4977-4980 vector form) are immediates: the stack of pending tree nodes lives in registers addressed by compile-time
// level, k is consumed in chunks of 8 so that the three lowest levels fold at compile time and only the levels above
// branch on the (warp-uniform) chunk counter.  Operands come through the read-only path: for a fixed k a warp reads
// 32 consecutive words of A (column-major) and one broadcast word of B.
#pragma once
#include "qb_device.cuh"

namespace qb {

// after leaf k: bit L of k set <=> a level-L node is pending in st[L] and merges with the new one
template <class Cfg, int L>
QB_DEV void tstack_push(typename Cfg::I (&st)[Cfg::NLEV + 1], uint32_t k, typename Cfg::I v)
{
    typedef typename Cfg::I I;
    if constexpr (L >= Cfg::NLEV) {
        st[L] = v;
    } else {
        if ((k >> L) & 1) {
            constexpr AddStage lv = Cfg::lvl(L);
            tstack_push<Cfg, L + 1>(st, k, merge_apply<I, true>(st[L], v, lv));
        } else {
            st[L] = v;
        }
    }
}

// the host-built flush program over the pending slots, unrolled at compile time
template <class Cfg, int IDX>
QB_DEV typename Cfg::I tstack_flush(const typename Cfg::I (&st)[Cfg::NLEV + 1], typename Cfg::I cur)
{
    typedef typename Cfg::I I;
    if constexpr (IDX >= Cfg::NFLUSH) {
        return cur;
    } else {
        constexpr FlushStep f = Cfg::flush(IDX);
        if constexpr (f.kind == FL_LOAD) cur = st[f.slot];
        else if constexpr (f.kind == FL_CONV) cur = conv_apply<I, true>(cur, f.a.q);
        else cur = add_apply<I, true>(st[f.slot], cur, f.a);
        return tstack_flush<Cfg, IDX + 1>(st, cur);
    }
}

template <class Cfg>
QB_DEV void gemul_thread_body(const typename Cfg::WA *__restrict__ A, const typename Cfg::WB *__restrict__ B,
                              typename Cfg::WC *__restrict__ C, uint64_t batch)
{
    typedef typename Cfg::I I;
    constexpr uint32_t M = Cfg::M, N = Cfg::N, K = Cfg::K, PER = M * N;
    constexpr QStage mulq = Cfg::mul(), outq = Cfg::out();
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= batch * PER) return;
    const uint64_t inst = t / PER;
    const uint32_t ij = (uint32_t)(t - inst * PER), j = ij / M, i = ij - j * M;
    const typename Cfg::WA *a = A + inst * (uint64_t)Cfg::SA;          // batch strides in words; 0 = shared operand
    const typename Cfg::WB *b = B + inst * (uint64_t)Cfg::SB;
    auto leaf = [&](uint32_t k) -> I {
        const uint32_t ia = Cfg::TRANS_A ? (k + i * K) : (i + k * M);                 // column-major, QuBLAS.h:2680-2692
        const uint32_t ib = Cfg::TRANS_B ? (j + k * N) : (k + j * K);
        I av = (I)__ldg(a + ia), bv = (I)__ldg(b + ib);
        if constexpr (Cfg::PRO_A) av = poly_ct<PolySel<Cfg, 1>, I>(av);      // fused ANUS prologue (qb200_gemul_prologue_batched)
        if constexpr (Cfg::PRO_B) bv = poly_ct<PolySel<Cfg, 2>, I>(bv);
        return mul_apply<I, true>(av, bv, mulq);
    };
    I st[Cfg::NLEV + 1];
    if constexpr (K <= 64) {
#pragma unroll
        for (uint32_t k = 0; k < K; ++k) tstack_push<Cfg, 0>(st, k, leaf(k));
    } else {
        constexpr uint32_t KC = K / 8;
#pragma unroll 4                                   // four chunks: the two lowest stack levels above the chunk fold as well
        for (uint32_t c = 0; c < KC; ++c) {
#pragma unroll
            for (uint32_t u = 0; u < 8; ++u) tstack_push<Cfg, 0>(st, c * 8 + u, leaf(c * 8 + u));
        }
#pragma unroll
        for (uint32_t k = KC * 8; k < K; ++k) tstack_push<Cfg, 0>(st, k, leaf(k));
    }
    I r;
    if constexpr (Cfg::NFLUSH == 0) r = st[Cfg::TOP_SLOT];
    else r = tstack_flush<Cfg, 0>(st, (I)0);
    I cv = conv_apply<I, true>(r, outq);
    if constexpr (Cfg::EPI) cv = poly_ct<Cfg, I>(cv);          // fused ANUS epilogue (qb200_gemul_poly_batched)
    C[inst * (uint64_t)Cfg::SC + ij] = (typename Cfg::WC)cv;
}

// ---------------------------------------------------------------------------------------------------------------
// Complex Qgemul, one thread per output element: TFComplexMul (QuBLAS.h:3522-3534), BasicComplexMul (:3437-3445) or
// real (x) complex (:3600-3637) products, two component-wise trees with their own level types (:3549-3568).
// Elements are {re, im} word pairs; a real operand (QB_CPROD_REAL_A / _B) stores one word per element.
// ---------------------------------------------------------------------------------------------------------------
template <typename W> QB_DEV void load_pair(const W *p, uint64_t e, W &re, W &im) { re = __ldg(p + 2 * e); im = __ldg(p + 2 * e + 1); }
template <> QB_DEV void load_pair<int32_t>(const int32_t *p, uint64_t e, int32_t &re, int32_t &im)
{
    const int2 v = __ldg((const int2 *)p + e);
    re = v.x; im = v.y;
}

template <class Cfg, int L>
QB_DEV void ctstack_push(typename Cfg::I (&sr)[Cfg::NLEV + 1], typename Cfg::I (&si)[Cfg::NLEV + 1], uint32_t k, typename Cfg::I vr, typename Cfg::I vi)
{
    typedef typename Cfg::I I;
    if constexpr (L >= Cfg::NLEV) {
        sr[L] = vr; si[L] = vi;
    } else {
        if ((k >> L) & 1) {
            constexpr AddStage lr = Cfg::lvl_re(L), li = Cfg::lvl_im(L);
            ctstack_push<Cfg, L + 1>(sr, si, k, merge_apply<I, true>(sr[L], vr, lr), merge_apply<I, true>(si[L], vi, li));
        } else {
            sr[L] = vr; si[L] = vi;
        }
    }
}

template <class Cfg, bool IM, int IDX>
QB_DEV typename Cfg::I ctstack_flush(const typename Cfg::I (&st)[Cfg::NLEV + 1], typename Cfg::I cur)
{
    typedef typename Cfg::I I;
    if constexpr (IDX >= (IM ? Cfg::NFLUSH_IM : Cfg::NFLUSH_RE)) {
        return cur;
    } else {
        constexpr FlushStep f = IM ? Cfg::flush_im(IDX) : Cfg::flush_re(IDX);
        if constexpr (f.kind == FL_LOAD) cur = st[f.slot];
        else if constexpr (f.kind == FL_CONV) cur = conv_apply<I, true>(cur, f.a.q);
        else cur = add_apply<I, true>(st[f.slot], cur, f.a);
        return ctstack_flush<Cfg, IM, IDX + 1>(st, cur);
    }
}

template <class Cfg>
QB_DEV void cgemul_thread_body(const typename Cfg::WA *__restrict__ A, const typename Cfg::WB *__restrict__ B,
                               typename Cfg::WC *__restrict__ C, uint64_t batch)
{
    typedef typename Cfg::I I;
    constexpr uint32_t M = Cfg::M, N = Cfg::N, K = Cfg::K, PER = M * N;
    constexpr int KIND = Cfg::KIND;
    constexpr TfProg tf = Cfg::tf();
    constexpr BcmProg bcm = Cfg::bcm();
    constexpr QStage ore = Cfg::out_re(), oim = Cfg::out_im();
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= batch * PER) return;
    const uint64_t inst = t / PER;
    const uint32_t ij = (uint32_t)(t - inst * PER), j = ij / M, i = ij - j * M;
    const uint64_t oa = inst * (uint64_t)Cfg::SA, ob = inst * (uint64_t)Cfg::SB;      // batch strides in elements; 0 = shared operand
    auto leaf = [&](uint32_t k, I &vr, I &vi) {
        const uint32_t ia = Cfg::TRANS_A ? (k + i * K) : (i + k * M);
        const uint32_t ib = Cfg::TRANS_B ? (j + k * N) : (k + j * K);
        typename Cfg::WA a = 0, b = 0;
        typename Cfg::WB c = 0, d = 0;
        if constexpr (KIND == 2) a = __ldg(A + oa + ia); else load_pair<typename Cfg::WA>(A, oa + ia, a, b);
        if constexpr (KIND == 3) c = __ldg(B + ob + ib); else load_pair<typename Cfg::WB>(B, ob + ib, c, d);
        cprod_apply<I, true>(KIND, (I)a, (I)b, (I)c, (I)d, tf, bcm, vr, vi);
    };
    I sr[Cfg::NLEV + 1], si[Cfg::NLEV + 1];
    if constexpr (K <= 32) {
#pragma unroll
        for (uint32_t k = 0; k < K; ++k) { I vr, vi; leaf(k, vr, vi); ctstack_push<Cfg, 0>(sr, si, k, vr, vi); }
    } else {
        constexpr uint32_t KC = K / 8;
#pragma unroll 4                                   // four chunks: the two lowest stack levels above the chunk fold as well
        for (uint32_t c = 0; c < KC; ++c) {
#pragma unroll
            for (uint32_t u = 0; u < 8; ++u) { I vr, vi; leaf(c * 8 + u, vr, vi); ctstack_push<Cfg, 0>(sr, si, c * 8 + u, vr, vi); }
        }
#pragma unroll
        for (uint32_t k = KC * 8; k < K; ++k) { I vr, vi; leaf(k, vr, vi); ctstack_push<Cfg, 0>(sr, si, k, vr, vi); }
    }
    I rr, ri;
    if constexpr (Cfg::NFLUSH_RE == 0) rr = sr[Cfg::TOP_RE]; else rr = ctstack_flush<Cfg, false, 0>(sr, (I)0);
    if constexpr (Cfg::NFLUSH_IM == 0) ri = si[Cfg::TOP_IM]; else ri = ctstack_flush<Cfg, true, 0>(si, (I)0);
    const uint64_t ic = inst * (uint64_t)Cfg::SC + ij;
    C[2 * ic] = (typename Cfg::WC)conv_apply<I, true>(rr, ore);
    C[2 * ic + 1] = (typename Cfg::WC)conv_apply<I, true>(ri, oim);
}

}  // namespace qb
