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
    const typename Cfg::WA *a = A + inst * (uint64_t)(M * K);
    const typename Cfg::WB *b = B + inst * (uint64_t)(K * N);
    auto leaf = [&](uint32_t k) -> I {
        const uint32_t ia = Cfg::TRANS_A ? (k + i * K) : (i + k * M);                 // column-major, QuBLAS.h:2680-2692
        const uint32_t ib = Cfg::TRANS_B ? (j + k * N) : (k + j * K);
        return mul_apply<I, true>((I)__ldg(a + ia), (I)__ldg(b + ib), mulq);
    };
    I st[Cfg::NLEV + 1];
    if constexpr (K <= 64) {
#pragma unroll
        for (uint32_t k = 0; k < K; ++k) tstack_push<Cfg, 0>(st, k, leaf(k));
    } else {
        constexpr uint32_t KC = K / 8;
#pragma unroll 1
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
    C[inst * (uint64_t)PER + ij] = (typename Cfg::WC)conv_apply<I, true>(r, outq);
}

}  // namespace qb
