// qb_fast_gemul.cuh — compile-time-specialised batched Qgemul for small dense instances
// (BASELINE config 3: 10^6 x 64x64x64, 16-bit operands) on the integer ALU.
//
//   * one CTA = one instance at a time, persistent over the batch (grid = resident CTAs x SMs);
//   * A and B instances (contiguous column-major arrays) are staged in shared memory by 1-D TMA
//     bulk copies (cp.async.bulk ... mbarrier::complete_tx) into a 2-stage ring, so the next
//     instance streams in from HBM while the current one is computed;
//   * each thread owns a TM x TN register tile of C and consumes k in order: 8 products are
//     reduced by a fully unrolled 3-level adjacent-pair tree, the resulting level-3 nodes go
//     through a level-indexed register stack (uniform branches on the chunk counter), which is
//     exactly the reference's pairwise order (QuBLAS.h:4936-4941) with its per-level quantizer;
//   * every shift / bound / mode of Qmul<MulArgs> (QuBLAS.h:3146-3171), of the per-level
//     Qadd (QuBLAS.h:3177-3204) and of the final converting assignment (QuBLAS.h:2398-2411) is an
//     immediate; clamps the host's interval analysis proved dead are not emitted.
#pragma once
#include "qb_device.cuh"

namespace qb {

QB_DEV uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

QB_DEV void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
QB_DEV void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
QB_DEV void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
QB_DEV void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <class Cfg, int TM, int TN>
struct GemulTile {
    typedef typename Cfg::I I;
    static constexpr int M = Cfg::M, N = Cfg::N, K = Cfg::K;
    static constexpr int KC = 8;                               // products reduced in registers per chunk
    static constexpr int NCH = K / KC;
    static constexpr int NTOP = Cfg::NLEV - 3;                 // stack levels above a chunk
    static constexpr int G = Cfg::G;                             // instances a CTA works on at a time (small shapes: fill the block)
    static constexpr int TPI = (M / TM) * (N / TN);              // threads per instance, each owns a 4x4 tile
    static constexpr int THREADS = G * TPI;                      // threads that own a tile
    static constexpr int BLOCK = (THREADS + 31) / 32 * 32;       // launched threads: the rest only helps with staging, conversion and barriers
    static constexpr int A_WORDS = M * K, B_WORDS = K * N;       // per instance
    // operand words: int32 (the reference's storage) or packed int16 (qb_qdesc.word_bytes = 2: half the HBM / PCIe bytes).
    // Packed words are staged as they are and widened by the conversion pass into a separate single-buffered region.
    static constexpr int WORD = (int)sizeof(typename Cfg::WA);
    static constexpr bool NARROW = WORD == 2;
    // One pass over each staged instance (gemul_tile_body): exact int->float conversion for the FP32-FMA product quantizer,
    // the per-instance operand ranges of the dynamic clamp elision, widening of packed words.
    // SUBST: the surviving clamp is min x min only and the conversion pass folds it into the operands (qb_device.cuh:fp_subst):
    // one clamp-free loop serves every instance.  Otherwise DYN: per-instance ranges pick the clamped or the clamp-free loop.
    static constexpr FpSubst SUB = fp_subst(Cfg::FP_MUL, Cfg::FP_SHIFT, Cfg::mul(), Cfg::A_LO, Cfg::A_HI, Cfg::B_LO, Cfg::B_HI);
    static constexpr bool SUBST = SUB.ok;
    static constexpr bool DYN = !SUBST && (Cfg::mul().need_lo || Cfg::mul().need_hi) && (Cfg::mul().of == SAT_TCPL || Cfg::mul().of == SAT_SMGN);
    static constexpr bool PRO = Cfg::PRO_A || Cfg::PRO_B;       // ANUS polynomial applied to an operand as it is staged (prologue)
    static constexpr bool PASS = DYN || Cfg::FP_MUL || NARROW || PRO;
    // SEP: the pass writes a separate single-buffered 32-bit copy (the host grants it when shared memory allows) instead of
    // converting in place; only then can the copy be laid out differently from the staged instance (swizzles below).
    static constexpr bool SEP = PASS && Cfg::SEP;
    static constexpr uint32_t STAGE_BYTES = G * (A_WORDS + B_WORDS) * WORD;
    static constexpr uint32_t WIDE_BYTES = SEP ? G * (A_WORDS + B_WORDS) * 4 : 0;
    // With the separate copy one staging buffer is enough: it is free again as soon as the pass is over, and the next
    // instance streams into it while the tiles work on the copy.  Converting in place needs the two-stage ring.
    static constexpr int NSTAGE = SEP ? 1 : 2;
    static constexpr uint32_t SMEM_BYTES = NSTAGE * STAGE_BYTES + WIDE_BYTES + 64;
    // Lane -> tile mapping.  The cost of a 16-byte shared-memory load depends on how the lanes share addresses
    // (tools/lds_probe.cu, profiles/r2_lds_probe.jsonl; ncu counts the same two classes as 4 and 2 wavefronts): it is at
    // the fast rate when every aligned group of four lanes reads at most two distinct 16-byte chunks, and takes 1.6x as long
    // otherwise -- no matter how few chunks the warp reads in total.  With tiles numbered row by row, four consecutive lanes
    // read four different A chunks.  PATCH: a warp owns 8 x 4 tiles with the lane bits interleaved (tx0, ty0, tx1, ty1, tx2),
    // so a group of four lanes spans 2 x 2 tiles: two A chunks and two B chunks; the warp reads 128 contiguous bytes of A.
    static constexpr bool PATCH = G == 1 && (M / TM) % 8 == 0 && (N / TN) % 4 == 0;
    // A K-contiguous operand (A stored transposed, B stored plain) puts the row blocks of a warp 4*K words apart, i.e. in the
    // same banks; the separate copy XORs the 16-byte chunk index with (row / 4) % 4, which spreads them over four bank groups.
    static constexpr bool SWZ_A = SEP && PATCH && Cfg::TRANS_A && K % 16 == 0;
    static constexpr bool SWZ_B = SEP && PATCH && !Cfg::TRANS_B && K % 16 == 0;
    static_assert(!NARROW || SEP, "packed words are widened into the separate copy");
    static_assert(M % TM == 0 && N % TN == 0 && TM == 4 && TN == 4, "tile shape");
    static constexpr bool K_POW2 = (K & (K - 1)) == 0;         // otherwise the pending nodes go through the flush program
    static_assert(K % KC == 0 && K >= KC, "K must be a multiple of 8");
    static_assert(sizeof(typename Cfg::WA) == sizeof(typename Cfg::WB) && (WORD == 4 || WORD == 2) && sizeof(typename Cfg::WC) == 4, "int32 or int16 operand words, int32 results");
    static_assert(!NARROW || (A_WORDS % 8 == 0 && B_WORDS % 8 == 0), "16-byte aligned int16 instances");
    // batch strides in words (plan.stride_*): dense by default, 0 = the operand is shared by all instances, larger = padded
    static constexpr unsigned long long SA = Cfg::SA, SB = Cfg::SB, SC = Cfg::SC;
    static_assert(G == 1 || (SA == (unsigned long long)A_WORDS && SB == (unsigned long long)B_WORDS), "groups of instances need dense operands");
    static_assert(SA % (16 / WORD) == 0 && SB % (16 / WORD) == 0 && SC % 4 == 0, "16-byte aligned instances");
};

// three-input min / max (one VIMNMX3 each on sm_90+)
QB_DEV int min3i(int a, int b, int c) { return __vimin3_s32(a, b, c); }
QB_DEV int max3i(int a, int b, int c) { return __vimax3_s32(a, b, c); }

// level-indexed register stack above the 8-product chunks; c = chunk counter (uniform)
template <class Cfg, int L, int NTOP, typename I>
QB_DEV void stack_push(I (&st)[NTOP > 0 ? NTOP : 1], uint32_t c, I v, I &result)
{
    if constexpr (L >= NTOP) {
        result = v;
    } else {
        if ((c >> L) & 1) {
            constexpr AddStage lv = Cfg::lvl(3 + L);
            I m = merge_apply<I, true>(st[L], v, lv);
            stack_push<Cfg, L + 1, NTOP, I>(st, c, m, result);
        } else {
            st[L] = v;
        }
    }
}

// Occupancy target: 512 resident threads per SM at the 128-register cap, bounded by shared memory.  Measured on B200
// (config 3, via QB200_JIT_DEFINES=-DQB_MIN_BLOCKS=n): 2 CTAs of 256 threads (16 warps, 52 words spilled) and 1 CTA with
// 202 registers and no spills run within noise of each other, 3 CTAs (85 registers) are 6 % slower, and prefetching the
// operands one half-chunk ahead (235 registers) changes nothing -- the loop is bound by pipe issue, not by latency.
constexpr int gemul_min_blocks(int threads, unsigned smem_bytes)
{
#ifdef QB_MIN_BLOCKS
    (void)threads; (void)smem_bytes;
    return QB_MIN_BLOCKS;
#else
    int n = 512 / threads, by_smem = (int)((227u * 1024u) / (smem_bytes + 1024u));
    if (n > by_smem) n = by_smem;
    return n < 1 ? 1 : (n > 8 ? 8 : n);
#endif
}

// K = 8 * (not a power of two): the same stack, but nodes stay pending at the end and are combined by the host-built
// flush program (vector form QuBLAS.h:4977-4980, variadic form :4942-4949), unrolled at compile time.  Slot = level - 3.
template <class Cfg, int L, int NTOP, typename I>
QB_DEV void stack_push_pending(I (&st)[NTOP > 0 ? NTOP : 1], uint32_t c, I v)
{
    if constexpr (L < NTOP) {
        if ((c >> L) & 1) {
            constexpr AddStage lv = Cfg::lvl(3 + L);
            stack_push_pending<Cfg, L + 1, NTOP, I>(st, c, merge_apply<I, true>(st[L], v, lv));
        } else {
            st[L] = v;
        }
    }
}
template <class Cfg, int IDX, int NTOP, typename I>
QB_DEV I stack_flush(const I (&st)[NTOP > 0 ? NTOP : 1], I cur)
{
    if constexpr (IDX >= Cfg::NFLUSH) {
        return cur;
    } else {
        constexpr FlushStep f = Cfg::flush(IDX);
        if constexpr (f.kind == FL_LOAD) cur = st[f.slot - 3];
        else if constexpr (f.kind == FL_CONV) cur = conv_apply<I, true>(cur, f.a.q);
        else cur = add_apply<I, true>(st[f.slot - 3], cur, f.a);
        return stack_flush<Cfg, IDX + 1, NTOP, I>(st, cur);
    }
}

constexpr QStage stage_without_clamps(QStage q) { q.need_lo = 0; q.need_hi = 0; return q; }

// ---- deferred-shift tree ("S form") -------------------------------------------------------------------------------
// When level 0 is an exact add and every level above it is a clamp-free halving with a data-independent bias --
// v = (x + y + b) >> 1 with b = 1 (RND::POS_INF) or b = 0 (RND::NEG_INF, TRN::TCPL at a 1-bit shift) -- a node is carried
// as its PRE-SHIFT sum S (value = S >> 1), and a merge is S' = (Sa >> 1) + (Sb >> 1) + b: one SHF and one LEA.HI.  The
// bias needs no instruction at all: ((S + 2e) >> 1) == (S >> 1) + e exactly, so the +b of a level-l node is pushed down
// its chain of left children and lands, as b * 2^(l-1), in the additive constant of the leftmost 4-product group below
// it -- a constant the level-0/1 IADD3 carries anyway (it also removes the FFMA magic-constant bias).  Same bits as the
// per-level quantizers (QuBLAS.h:3185-3194 applied per tree level, :4936-4941); tests compare with the oracle and with
// the runtime-descriptor kernels.
template <class Cfg> __host__ __device__ constexpr bool s_form_ok()
{
    typedef GemulTile<Cfg, 4, 4> T;
#ifndef QB_SFORM_MAX_NCH
#define QB_SFORM_MAX_NCH 16                                     // K <= 128: 64x64x128 runs at 10.3 TMAC/s in this form, 9.0 without
#endif
    if (!T::K_POW2 || Cfg::NLEV < 3 || T::NCH > QB_SFORM_MAX_NCH || T::BLOCK > 512) return false;       // needs the fully unrolled chunk loop
    constexpr AddStage l0 = Cfg::lvl(0);
    if (l0.q.shift != 0 || l0.q.need_lo || l0.q.need_hi || l0.q.wide || l0.shl_a || l0.shl_b || l0.sub) return false;
    if (l0.q.of == WRP_TCPL && (l0.q.need_lo || l0.q.need_hi)) return false;
    for (int l = 1; l < Cfg::NLEV; ++l) {
        const AddStage s = Cfg::lvl(l);
        if (s.q.shift != 1 || s.q.need_lo || s.q.need_hi || s.q.wide || s.shl_a || s.shl_b || s.sub) return false;
        if (!(s.q.qu == RND_POS_INF || s.q.qu == RND_NEG_INF || s.q.qu == TRN_TCPL)) return false;
    }
    return true;
}
// Negated S form (round 2).  ceil(x / 2) = -floor(-x / 2): if every node is carried NEGATED, a RND::POS_INF halving needs no
// bias at all (and a round-down level needs the +1 instead).  The negation is free -- the conversion pass negates the float
// copy of A, RND::CONV products are odd-symmetric -- and with level 1 rounding half-up the additive constant of every group
// but the few that inherit a bias from a round-down level above is ZERO.  The FFMA magic-constant bias (4 * fp_kbits per
// group, an even number at every level: ((S + D) >> 1) + ((S' + D) >> 1) = (S >> 1) + (S' >> 1) + D) is not removed per group
// any more but once per output at the root.  A group is then IADD3(p0, p1, p2) + one plain two-input add without an
// immediate (config 3: 11.4 -> 11.9 TMAC/s).  Issuing that add as IMAD (x * 1 + y, the 1 in a register) on the FMA pipe gains
// 5 % from registers (tools/mix_probe.cu) but loses 1 % in the kernel (QB_IMAD_ADD).
template <class Cfg> __host__ __device__ constexpr bool s_form_neg()
{
#ifdef QB_NO_NEG
    return false;
#else
    return GemulTile<Cfg, 4, 4>::SUBST && s_form_ok<Cfg>() && Cfg::lvl(1).q.qu == RND_POS_INF;
#endif
}
// additive constant of 4-product group g (0 .. K/4-1): the level-1 bias plus every bias pushed down from the levels above
template <class Cfg> __host__ __device__ constexpr int s_form_const(int g)
{
    const bool neg = s_form_neg<Cfg>();
    int c = ((Cfg::lvl(1).q.qu == RND_POS_INF) != neg) ? 1 : 0;
    for (int l = 2; l < Cfg::NLEV; ++l)
        if (g % (1 << (l - 1)) == 0 && ((Cfg::lvl(l).q.qu == RND_POS_INF) != neg)) c += 1 << (l - 1);
    return c;
}
template <class Cfg> __host__ __device__ constexpr bool s_form_no_consts()
{
    for (int g = 0; g < Cfg::K / 4; ++g) if (s_form_const<Cfg>(g) != 0) return false;
    return true;
}
#ifdef QB_IMAD_ADD
// the 1 of the IMAD adds: read from memory so that the compiler cannot fold the multiply away
static __device__ uint32_t qb_opaque_one = 1u;
#endif
// level-indexed stack of pre-shift sums; slot L holds a pending node of level 2 + L
template <class Cfg, int L, int NTOP, typename I>
QB_DEV void s_push(I (&st)[NTOP > 0 ? NTOP : 1], uint32_t c, I S, I &result)
{
    if constexpr (L >= NTOP) {
        result = S;
    } else {
        if ((c >> L) & 1) s_push<Cfg, L + 1, NTOP, I>(st, c, (I)((st[L] >> 1) + (S >> 1)), result);
        else st[L] = S;
    }
}

// Fragment addressing shared by both tile loops: 16-byte chunk pointers into the 32-bit operands the tiles read.  kc = k0 / 4
// is a compile-time constant once the chunk loop is unrolled, so with a swizzled operand the address is one of four
// per-thread base registers (kc % 4 XOR the thread's row-block index) plus an immediate.
template <class Cfg>
struct TileFrag {
    typedef GemulTile<Cfg, 4, 4> T;
    const int4 *a[4], *b[4];
    QB_DEV TileFrag(const int32_t *As, const int32_t *Bs, int i0, int j0)
    {
        const int swa = T::SWZ_A ? (i0 >> 2) & 3 : 0, swb = T::SWZ_B ? (j0 >> 2) & 3 : 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            a[q] = (const int4 *)As + (Cfg::TRANS_A ? i0 * (T::K / 4) + (q ^ swa) : (i0 >> 2));
            b[q] = (const int4 *)Bs + (Cfg::TRANS_B ? (j0 >> 2) : j0 * (T::K / 4) + (q ^ swb));
        }
    }
    // A: row i0 + u at k-chunk kc (stored transposed: k contiguous) or rows i0 .. i0+3 at k = 4 kc + u (plain: i contiguous)
    QB_DEV int4 lda(int u, int kc) const { return Cfg::TRANS_A ? a[kc & 3][u * (T::K / 4) + (kc & ~3)] : a[0][((4 * kc + u) * T::M) >> 2]; }
    // B: column j0 + u at k-chunk kc (plain: k contiguous) or columns j0 .. j0+3 at k = 4 kc + u (stored transposed)
    QB_DEV int4 ldb(int u, int kc) const { return Cfg::TRANS_B ? b[0][((4 * kc + u) * T::N) >> 2] : b[kc & 3][u * (T::K / 4) + (kc & ~3)]; }
};

// One instance's 4x4 register tile.  CLAMP = false is the same arithmetic with Qmul's saturation removed;
// it is selected per instance when the instance's own operand ranges prove that no product can saturate.
template <class Cfg, bool CLAMP>
QB_DEV void tile_compute(const int32_t *As, const int32_t *Bs, int i0, int j0, typename Cfg::I (&res)[4][4])
{
    typedef GemulTile<Cfg, 4, 4> T;
    typedef typename T::I I;
    constexpr int NTOP = T::NTOP;
    constexpr QStage mulq = CLAMP ? Cfg::mul() : stage_without_clamps(Cfg::mul());
    constexpr AddStage l0 = Cfg::lvl(0), l1 = Cfg::lvl(1), l2 = Cfg::lvl(2);
    const TileFrag<Cfg> fr(As, Bs, i0, j0);
    I st[4][4][NTOP > 0 ? NTOP : 1];
    // The chunk loop is unrolled (up to 8 chunks = the whole K = 64 of config 3): the branches of the level-indexed stack
    // on the chunk counter then fold at compile time, the stack slots get exact live ranges instead of being loop-carried,
    // and the scheduler can overlap the tail of one chunk with the loads of the next.  Measured on config 3 (NVRTC build):
    // unroll 1 / 2 / 4 / 8 = 8 398 / 8 978 / 8 929 / 10 244 GMAC/s.  QB_CHUNK_UNROLL overrides (kernel experiments).
#ifndef QB_CHUNK_UNROLL
#define QB_CHUNK_UNROLL 0
#endif
    // (not for blocks above 512 threads: their register cap is 64 and the unrolled body spills: 128x128x64 drops from 8.3 to 6.7 TMAC/s)
    constexpr int CHUNK_UNROLL = QB_CHUNK_UNROLL > 0 ? QB_CHUNK_UNROLL : (T::BLOCK > 512 ? 1 : (T::NCH <= 8 ? T::NCH : 8));
#pragma unroll CHUNK_UNROLL
    for (uint32_t c = 0; c < (uint32_t)T::NCH; ++c) {
        I half0[4][4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            // a 4 (i) x 4 (k) block of op(A) and a 4 (k) x 4 (j) block of op(B) as four 16-byte loads each; which index runs
            // inside a load depends on the storage: A is M x K column-major (i contiguous), transposed K x M (k contiguous);
            // B is K x N column-major (k contiguous), transposed N x K (j contiguous)
            int4 bq[4], aq[4];
            if constexpr (CHUNK_UNROLL >= T::NCH) {
#pragma unroll
                for (int u = 0; u < 4; ++u) { bq[u] = fr.ldb(u, 2 * (int)c + h); aq[u] = fr.lda(u, 2 * (int)c + h); }
            } else {
                // rolled chunk loop: the chunk index is a run-time value, the XOR of a swizzled operand is computed per load
                const int k0 = c * 8 + h * 4;
                const int swa = T::SWZ_A ? (i0 >> 2) & 3 : 0, swb = T::SWZ_B ? (j0 >> 2) & 3 : 0;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    bq[u] = Cfg::TRANS_B ? ((const int4 *)Bs)[((k0 + u) * T::N + j0) >> 2] : ((const int4 *)Bs)[(j0 + u) * (T::K / 4) + ((k0 >> 2) ^ swb)];
                    aq[u] = Cfg::TRANS_A ? ((const int4 *)As)[(i0 + u) * (T::K / 4) + ((k0 >> 2) ^ swa)] : ((const int4 *)As)[((k0 + u) * T::M + i0) >> 2];
                }
            }
            auto a_at = [&](int ii, int kk) -> I { return Cfg::TRANS_A ? (&aq[ii].x)[kk] : (&aq[kk].x)[ii]; };
            auto b_at = [&](int kk, int jj) -> I { return Cfg::TRANS_B ? (&bq[kk].x)[jj] : (&bq[jj].x)[kk]; };
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const I b0 = b_at(0, jj), b1 = b_at(1, jj), b2 = b_at(2, jj), b3 = b_at(3, jj);
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) {
                    const I a0 = a_at(ii, 0), a1 = a_at(ii, 1), a2 = a_at(ii, 2), a3 = a_at(ii, 3);
                    I s01, s23;
                    if constexpr (Cfg::FP_MUL) {
                        constexpr int S = Cfg::FP_SHIFT;
                        constexpr uint32_t KB2 = 2u * (uint32_t)fp_kbits<S>();
                        const int p0 = fp_mul_apply<S, true>(__int_as_float(a0), __int_as_float(b0), mulq);
                        const int p1 = fp_mul_apply<S, true>(__int_as_float(a1), __int_as_float(b1), mulq);
                        const int p2 = fp_mul_apply<S, true>(__int_as_float(a2), __int_as_float(b2), mulq);
                        const int p3 = fp_mul_apply<S, true>(__int_as_float(a3), __int_as_float(b3), mulq);
                        s01 = q_apply<I, true>((I)((uint32_t)p0 + (uint32_t)p1 - KB2), l0.q);       // unbias folded into the level-0 add
                        s23 = q_apply<I, true>((I)((uint32_t)p2 + (uint32_t)p3 - KB2), l0.q);
                    } else {
                        I p0 = mul_apply<I, true>(a0, b0, mulq);
                        I p1 = mul_apply<I, true>(a1, b1, mulq);
                        I p2 = mul_apply<I, true>(a2, b2, mulq);
                        I p3 = mul_apply<I, true>(a3, b3, mulq);
                        s01 = merge_apply<I, true>(p0, p1, l0);
                        s23 = merge_apply<I, true>(p2, p3, l0);
                    }
                    const I hv = merge_apply<I, true>(s01, s23, l1);
                    if (h == 0) {
                        half0[ii][jj] = hv;
                    } else {
                        const I v = merge_apply<I, true>(half0[ii][jj], hv, l2);
                        if constexpr (T::K_POW2) stack_push<Cfg, 0, NTOP, I>(st[ii][jj], c, v, res[ii][jj]);
                        else stack_push_pending<Cfg, 0, NTOP, I>(st[ii][jj], c, v);
                    }
                }
            }
        }
    }
    if constexpr (!T::K_POW2) {
#pragma unroll
        for (int ii = 0; ii < 4; ++ii)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) res[ii][jj] = stack_flush<Cfg, 0, NTOP, I>(st[ii][jj], (I)0);
    }
}

// The same tile in S form (s_form_ok): JW of its four columns per pass over K (JW = 4: one pass; JW = 2: two passes, the
// A fragments are read twice from shared memory but only 8 accumulator chains are live).  out(jj, v0..v3) receives the four
// finished rows of column j0 + jj.
#ifndef QB_SFORM_JW
#define QB_SFORM_JW 2
#endif
template <class Cfg, bool CLAMP, int JW, class Out>
QB_DEV void tile_compute_s(const int32_t *As, const int32_t *Bs, int i0, int j0, uint32_t one, Out out)
{
    constexpr bool NEG = s_form_neg<Cfg>();
    static_assert(!NEG || (!CLAMP && Cfg::FP_MUL), "the negated form runs on substituted operands");
    typedef GemulTile<Cfg, 4, 4> T;
    typedef typename T::I I;
    constexpr int NTOP = T::NTOP;
    constexpr QStage mulq = CLAMP ? Cfg::mul() : stage_without_clamps(Cfg::mul());
    static_assert(JW == 4 || JW == 2, "columns per pass");
    // the passes must stay sequential (unrolled, the scheduler interleaves them and every chain is live again)
#pragma unroll 1
    for (int jp = 0; jp < 4; jp += JW) {
        asm volatile("" ::: "memory");                        // keep the (pass-invariant) A fragment loads inside their pass
        const TileFrag<Cfg> fr(As, Bs + (Cfg::TRANS_B ? 0 : jp * T::K), i0, j0);      // plain B: the pass starts at column j0 + jp (same row block: same swizzle)
        I st[4][JW][NTOP > 0 ? NTOP : 1], res[4][JW];
#ifdef QB_SFORM_PREFETCH                                      // experiment: fragments of the next four k in flight behind the current group
        int4 bqn[4], aqn[4];
#pragma unroll
        for (int u = 0; u < (Cfg::TRANS_B ? 4 : JW); ++u) bqn[u] = fr.ldb(u, 0);
#pragma unroll
        for (int u = 0; u < 4; ++u) aqn[u] = fr.lda(u, 0);
#endif
#pragma unroll
        for (uint32_t c = 0; c < (uint32_t)T::NCH; ++c) {
            I half0[4][JW];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int kc = 2 * (int)c + h;
                int4 bq[4], aq[4];
#ifdef QB_SFORM_PREFETCH
#pragma unroll
                for (int u = 0; u < 4; ++u) { bq[u] = bqn[u]; aq[u] = aqn[u]; }
                if (kc + 1 < 2 * T::NCH) {
#pragma unroll
                    for (int u = 0; u < (Cfg::TRANS_B ? 4 : JW); ++u) bqn[u] = fr.ldb(u, kc + 1);
#pragma unroll
                    for (int u = 0; u < 4; ++u) aqn[u] = fr.lda(u, kc + 1);
                }
#else
#pragma unroll
                for (int u = 0; u < (Cfg::TRANS_B ? 4 : JW); ++u) bq[u] = fr.ldb(u, kc);            // stored transposed: j contiguous, all four columns come along
#pragma unroll
                for (int u = 0; u < 4; ++u) aq[u] = fr.lda(u, kc);
#endif
                auto a_at = [&](int ii, int kk) -> I { return Cfg::TRANS_A ? (&aq[ii].x)[kk] : (&aq[kk].x)[ii]; };
                auto b_at = [&](int kk, int jj) -> I {
                    if constexpr (!Cfg::TRANS_B) return (&bq[jj].x)[kk];
                    else if constexpr (JW == 4) return (&bq[kk].x)[jj];
                    else return jp ? (&bq[kk].z)[jj] : (&bq[kk].x)[jj];            // second pass: columns 2, 3
                };
#pragma unroll
                for (int jj = 0; jj < JW; ++jj) {
                    const I b0 = b_at(0, jj), b1 = b_at(1, jj), b2 = b_at(2, jj), b3 = b_at(3, jj);
#pragma unroll
                    for (int ii = 0; ii < 4; ++ii) {
                        const I a0 = a_at(ii, 0), a1 = a_at(ii, 1), a2 = a_at(ii, 2), a3 = a_at(ii, 3);
                        // four products -> the pre-shift level-1 sum (two IADD3): level 0 is exact, the constant removes the
                        // FFMA bias and carries every rounding bias that belongs to this group
                        uint32_t sum;
                        if constexpr (NEG) {
                            // negated products, biased by fp_kbits each (removed at the root); level 0 is exact
                            constexpr int S = Cfg::FP_SHIFT;
                            const int KG = s_form_const<Cfg>(2 * (int)c + h);      // a constant once the chunk loop is unrolled
                            const uint32_t p3 = (uint32_t)fp_mul_apply<S, true>(__int_as_float(a3), __int_as_float(b3), mulq);
                            sum = (uint32_t)fp_mul_apply<S, true>(__int_as_float(a0), __int_as_float(b0), mulq) + (uint32_t)fp_mul_apply<S, true>(__int_as_float(a1), __int_as_float(b1), mulq) +
                                  (uint32_t)fp_mul_apply<S, true>(__int_as_float(a2), __int_as_float(b2), mulq);
#ifdef QB_IMAD_ADD                                            // experiment: the plain add as IMAD (x * 1 + y) on the FMA pipe; measured -1 % in the kernel
                            if constexpr (s_form_no_consts<Cfg>()) sum = sum * one + p3;
                            else sum = sum + p3 + (uint32_t)KG;
#else
                            sum = sum + p3 + (uint32_t)KG;
#endif
                            const I S1 = (I)sum;
                            if (h == 0) half0[ii][jj] = S1;
                            else s_push<Cfg, 0, NTOP, I>(st[ii][jj], c, (I)((half0[ii][jj] >> 1) + (S1 >> 1)), res[ii][jj]);
                            continue;
                        } else if constexpr (Cfg::FP_MUL) {
                            constexpr int S = Cfg::FP_SHIFT;
                            sum = (uint32_t)fp_mul_apply<S, true>(__int_as_float(a0), __int_as_float(b0), mulq) + (uint32_t)fp_mul_apply<S, true>(__int_as_float(a1), __int_as_float(b1), mulq) +
                                  (uint32_t)fp_mul_apply<S, true>(__int_as_float(a2), __int_as_float(b2), mulq) + (uint32_t)fp_mul_apply<S, true>(__int_as_float(a3), __int_as_float(b3), mulq) -
                                  4u * (uint32_t)fp_kbits<S>();
                        } else {
                            sum = (uint32_t)mul_apply<I, true>(a0, b0, mulq) + (uint32_t)mul_apply<I, true>(a1, b1, mulq) + (uint32_t)mul_apply<I, true>(a2, b2, mulq) + (uint32_t)mul_apply<I, true>(a3, b3, mulq);
                        }
                        const I S1 = (I)(sum + (uint32_t)s_form_const<Cfg>(2 * (int)c + h));
                        if (h == 0) half0[ii][jj] = S1;
                        else s_push<Cfg, 0, NTOP, I>(st[ii][jj], c, (I)((half0[ii][jj] >> 1) + (S1 >> 1)), res[ii][jj]);
                    }
                }
            }
        }
#pragma unroll
        for (int jj = 0; jj < JW; ++jj) {
            if constexpr (NEG) {                              // root: offset, negated pre-shift sum -> value
                constexpr I DH = (I)((int32_t)(4u * (uint32_t)fp_kbits<Cfg::FP_SHIFT>()) >> 1);
                out(jp + jj, (I)(DH - (res[0][jj] >> 1)), (I)(DH - (res[1][jj] >> 1)), (I)(DH - (res[2][jj] >> 1)), (I)(DH - (res[3][jj] >> 1)));
            } else {
                out(jp + jj, (I)(res[0][jj] >> 1), (I)(res[1][jj] >> 1), (I)(res[2][jj] >> 1), (I)(res[3][jj] >> 1));   // root: pre-shift sum -> value
            }
        }
    }
}

// Which operand extremes the dynamic clamp elision has to find per instance.  A corner product of the operand TYPES' own
// bounds that already satisfies the thresholds can never be the reason for a clamp, so the pass only tracks the side
// (minima, maxima) that takes part in a corner that can; the other side enters the check with the type's static bound.
template <class Cfg> __host__ __device__ constexpr bool dyn_violates(long long p)
{
    return (Cfg::mul().need_hi && p > Cfg::P_HI) || (Cfg::mul().need_lo && p < Cfg::P_LO);
}
template <class Cfg> __host__ __device__ constexpr bool dyn_track_hi()
{
    return dyn_violates<Cfg>(Cfg::A_LO * Cfg::B_HI) || dyn_violates<Cfg>(Cfg::A_HI * Cfg::B_LO) || dyn_violates<Cfg>(Cfg::A_HI * Cfg::B_HI);
}
template <class Cfg> __host__ __device__ constexpr bool dyn_track_lo()
{
    return dyn_violates<Cfg>(Cfg::A_LO * Cfg::B_HI) || dyn_violates<Cfg>(Cfg::A_HI * Cfg::B_LO) || dyn_violates<Cfg>(Cfg::A_LO * Cfg::B_LO);
}

template <class Cfg>
QB_DEV void gemul_tile_body(const typename Cfg::WA *__restrict__ A, const typename Cfg::WB *__restrict__ B, int32_t *__restrict__ C, uint64_t batch)
{
    typedef GemulTile<Cfg, 4, 4> T;
    typedef typename T::I I;
    constexpr int M = T::M;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bars = (uint64_t *)(smem_raw + T::NSTAGE * T::STAGE_BYTES + T::WIDE_BYTES);
    __shared__ int4 range_smem[32];                                // per warp: { min A, max A, min B, max B }

    const int tid = threadIdx.x;
    const int g = T::G == 1 ? 0 : tid / T::TPI, tl = T::G == 1 ? tid : tid - g * T::TPI;       // instance within the group, thread within the instance
    int tx, ty;
    if constexpr (T::PATCH) {                                      // a warp owns 8 x 4 tiles, lane bits interleaved (GemulTile::PATCH)
        constexpr int WX = (M / 4) / 8;
        const int wp = tl >> 5;
        tx = (wp % WX) * 8 + ((tl & 1) | ((tl >> 1) & 2) | ((tl >> 2) & 4));
        ty = (wp / WX) * 4 + (((tl >> 1) & 1) | ((tl >> 2) & 2));
    } else {
        tx = tl % (M / 4); ty = tl / (M / 4);
    }
    const int i0 = tx * 4, j0 = ty * 4;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // a work item is a group of G consecutive instances (the last group of the batch may be shorter); their A arrays are
    // contiguous in memory, and so are their B arrays: two bulk copies per group
    const uint64_t ngroups = (batch + T::G - 1) / T::G;
    auto group_count = [&](uint64_t grp) -> uint32_t {
        const uint64_t left = batch - grp * T::G;
        return (uint32_t)(left < (uint64_t)T::G ? left : (uint64_t)T::G);
    };
    auto issue = [&](uint64_t grp, int s) {
        unsigned char *dst = smem_raw + (size_t)s * T::STAGE_BYTES;
        const uint32_t cnt = group_count(grp);
        mbar_expect_tx(&bars[s], cnt * (uint32_t)(T::A_WORDS + T::B_WORDS) * (uint32_t)T::WORD);
        tma_load_1d(dst, A + grp * T::G * T::SA, cnt * (uint32_t)T::A_WORDS * (uint32_t)T::WORD, &bars[s]);
        tma_load_1d(dst + T::G * T::A_WORDS * T::WORD, B + grp * T::G * T::SB, cnt * (uint32_t)T::B_WORDS * (uint32_t)T::WORD, &bars[s]);
    };

    uint64_t inst = blockIdx.x;                                   // group index
    if (inst < ngroups && tid == 0) issue(inst, 0);

    constexpr QStage mulq = Cfg::mul();
    constexpr QStage outq = Cfg::out();
#ifdef QB_IMAD_ADD
    const uint32_t one = *(const volatile uint32_t *)&qb_opaque_one;
#else
    const uint32_t one = 1u;
#endif
    // Dynamic clamp elision: Qmul's SAT clamp survives the static interval analysis only because of extreme
    // operand values (config 3: min x min).  One pass over the staged instance finds its actual operand
    // ranges; if even the extreme corner products stay inside [P_LO, P_HI] (host-computed, conservative),
    // the clamp-free instantiation runs.  Same bits either way.
    constexpr bool DYN = T::DYN, PASS = T::PASS;
    constexpr bool TRACK_LO = DYN && dyn_track_lo<Cfg>(), TRACK_HI = DYN && dyn_track_hi<Cfg>();

    for (uint32_t it = 0; inst < ngroups; inst += gridDim.x, ++it) {
        const int s = T::NSTAGE == 2 ? it & 1 : 0;
        const uint64_t next = inst + gridDim.x;
        if constexpr (T::NSTAGE == 2) {
            if (tid == 0 && next < ngroups) {
                if constexpr (PASS && (Cfg::FP_MUL || T::PRO)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // in-place int->float / prologue writes (generic proxy) before the TMA refill
                issue(next, s ^ 1);                               // stage s^1 was released by the barrier that ended the previous iteration
            }
        }
        mbar_wait(&bars[s], T::NSTAGE == 2 ? (it >> 1) & 1 : it & 1);
        unsigned char *stage = smem_raw + (size_t)s * T::STAGE_BYTES;
        int32_t *As = T::SEP ? (int32_t *)(smem_raw + T::NSTAGE * T::STAGE_BYTES) : (int32_t *)stage;      // 32-bit operands the tiles read
        int32_t *Bs = As + T::G * T::A_WORDS;
        bool clamp = true;
        if constexpr (PASS) {
            // one pass over the staged operands: exact int->float conversion (FP_MUL), per-instance ranges (DYN), widening of
            // packed words, and the swizzle of K-contiguous operands (SEP: into the separate copy)
            int lo[2] = { 0x7fffffff, 0x7fffffff }, hi[2] = { (int)0x80000000, (int)0x80000000 };
            auto scan4 = [&](int part, const int4 &v) {
                if constexpr (TRACK_LO) lo[part] = min3i(min3i(lo[part], v.x, v.y), v.z, v.w);      // two VIMNMX3 per four words
                if constexpr (TRACK_HI) hi[part] = max3i(max3i(hi[part], v.x, v.y), v.z, v.w);
            };
            auto pro4 = [&](int part, int4 &v) {                                 // fused ANUS prologue on the raw words of an operand
                if constexpr (Cfg::PRO_A) if (part == 0) {
                    v.x = (int)poly_ct<PolySel<Cfg, 1>, I>((I)v.x); v.y = (int)poly_ct<PolySel<Cfg, 1>, I>((I)v.y);
                    v.z = (int)poly_ct<PolySel<Cfg, 1>, I>((I)v.z); v.w = (int)poly_ct<PolySel<Cfg, 1>, I>((I)v.w);
                }
                if constexpr (Cfg::PRO_B) if (part == 1) {
                    v.x = (int)poly_ct<PolySel<Cfg, 2>, I>((I)v.x); v.y = (int)poly_ct<PolySel<Cfg, 2>, I>((I)v.y);
                    v.z = (int)poly_ct<PolySel<Cfg, 2>, I>((I)v.z); v.w = (int)poly_ct<PolySel<Cfg, 2>, I>((I)v.w);
                }
            };
            auto put4 = [&](int part, int4 *w, int i, const int4 &v) {          // i = 16-byte chunk index within the part
                const bool swz = part ? T::SWZ_B : T::SWZ_A;
                const int o = swz ? i ^ ((i / T::K) & 3) : i;                    // chunk ^ (row / 4) % 4, row = i / (K / 4)
                if constexpr (T::SUBST) {                                        // an operand at its type's minimum -> the substitute that saturates min x min
                    const float sub = part ? T::SUB.b_sub : T::SUB.a_sub;
                    if (s_form_neg<Cfg>() && part == 0)                           // negated S form: A enters negated (FMNMX negates its inputs for free)
                        ((float4 *)w)[o] = make_float4(fminf(-__int2float_rn(v.x), -sub), fminf(-__int2float_rn(v.y), -sub), fminf(-__int2float_rn(v.z), -sub), fminf(-__int2float_rn(v.w), -sub));
                    else
                        ((float4 *)w)[o] = make_float4(fmaxf(__int2float_rn(v.x), sub), fmaxf(__int2float_rn(v.y), sub), fmaxf(__int2float_rn(v.z), sub), fmaxf(__int2float_rn(v.w), sub));
                } else if constexpr (Cfg::FP_MUL) ((float4 *)w)[o] = make_float4(__int2float_rn(v.x), __int2float_rn(v.y), __int2float_rn(v.z), __int2float_rn(v.w));
                else w[o] = v;
            };
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                int4 *w = (int4 *)(part ? Bs : As);
                if constexpr (T::NARROW) {
                    // packed int16 words: widen (sign-extend) into the 32-bit copy, converting / scanning on the way
                    const int4 *raw = (const int4 *)(stage + (part ? T::G * T::A_WORDS * 2 : 0));
                    const int n8 = (part ? T::G * T::B_WORDS : T::G * T::A_WORDS) / 8;
                    for (int i = tid; i < n8; i += T::BLOCK) {
                        const int4 v = raw[i];
                        int4 e0 = make_int4((int)(short)(v.x & 0xffff), v.x >> 16, (int)(short)(v.y & 0xffff), v.y >> 16);
                        int4 e1 = make_int4((int)(short)(v.z & 0xffff), v.z >> 16, (int)(short)(v.w & 0xffff), v.w >> 16);
                        pro4(part, e0); pro4(part, e1);
                        scan4(part, e0); scan4(part, e1);
                        put4(part, w, 2 * i, e0); put4(part, w, 2 * i + 1, e1);
                    }
                } else {
                    const int4 *raw = (const int4 *)(stage + (part ? T::G * T::A_WORDS * 4 : 0));
                    constexpr int N4A = T::G * T::A_WORDS / 4, N4B = T::G * T::B_WORDS / 4;      // the whole group (a short last group leaves stale words: harmless, conservative for the ranges)
                    const int n4 = part ? N4B : N4A;
                    constexpr int PER = ((N4A > N4B ? N4A : N4B) + T::BLOCK - 1) / T::BLOCK;
                    if constexpr (PER <= 8) {
                        // all loads of this part in flight first, then ranges / conversion / stores
                        int4 v[PER];
#pragma unroll
                        for (int u = 0; u < PER; ++u) { const int i = tid + u * T::BLOCK; if (i < n4) v[u] = raw[i]; }
#pragma unroll
                        for (int u = 0; u < PER; ++u) {
                            const int i = tid + u * T::BLOCK;
                            if (i < n4) { pro4(part, v[u]); scan4(part, v[u]); if constexpr (T::SEP || Cfg::FP_MUL || T::PRO) put4(part, w, i, v[u]); }
                        }
                    } else {
                        for (int i = tid; i < n4; i += T::BLOCK) {
                            int4 v = raw[i];
                            pro4(part, v);
                            scan4(part, v);
                            if constexpr (T::SEP || Cfg::FP_MUL || T::PRO) put4(part, w, i, v);
                        }
                    }
                }
            }
            if constexpr (DYN) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    if constexpr (TRACK_LO) { lo[0] = imin<int>(lo[0], __shfl_xor_sync(0xffffffffu, lo[0], o)); lo[1] = imin<int>(lo[1], __shfl_xor_sync(0xffffffffu, lo[1], o)); }
                    if constexpr (TRACK_HI) { hi[0] = imax<int>(hi[0], __shfl_xor_sync(0xffffffffu, hi[0], o)); hi[1] = imax<int>(hi[1], __shfl_xor_sync(0xffffffffu, hi[1], o)); }
                }
                if ((tid & 31) == 0) range_smem[tid >> 5] = make_int4(lo[0], hi[0], lo[1], hi[1]);
            }
            __syncthreads();
            if constexpr (T::NSTAGE == 1) {
                if (tid == 0 && next < ngroups) issue(next, 0);    // the pass has consumed the staged instance: stream the next one in behind the tiles
            }
            if constexpr (DYN) {
                int alo = 0x7fffffff, ahi = (int)0x80000000, blo = 0x7fffffff, bhi = (int)0x80000000;
#pragma unroll
                for (int wv = 0; wv < T::BLOCK / 32; wv += 2) {
                    const int4 r0 = range_smem[wv], r1 = wv + 1 < T::BLOCK / 32 ? range_smem[wv + 1] : r0;
                    if constexpr (TRACK_LO) { alo = min3i(alo, r0.x, r1.x); blo = min3i(blo, r0.z, r1.z); }
                    if constexpr (TRACK_HI) { ahi = max3i(ahi, r0.y, r1.y); bhi = max3i(bhi, r0.w, r1.w); }
                }
                if constexpr (!TRACK_LO) { alo = (int)Cfg::A_LO; blo = (int)Cfg::B_LO; }          // statically harmless side: the type's own bound
                if constexpr (!TRACK_HI) { ahi = (int)Cfg::A_HI; bhi = (int)Cfg::B_HI; }
                const long long c0 = (long long)alo * blo, c1 = (long long)alo * bhi, c2 = (long long)ahi * blo, c3 = (long long)ahi * bhi;
                const long long pmax = imax<long long>(imax<long long>(c0, c1), imax<long long>(c2, c3));
                const long long pmin = imin<long long>(imin<long long>(c0, c1), imin<long long>(c2, c3));
                clamp = (mulq.need_hi && pmax > Cfg::P_HI) || (mulq.need_lo && pmin < Cfg::P_LO);
            }
        }

        if ((T::BLOCK == T::THREADS || tid < T::THREADS) && (T::G == 1 || (uint32_t)g < group_count(inst))) {   // padding threads / a short last group own no tile
            const int32_t *Ag = As + g * T::A_WORDS, *Bg = Bs + g * T::B_WORDS;
            // converting assignment into C's element type, column-major 128-bit stores
            int32_t *Ci = C + (inst * T::G + g) * T::SC;
            auto store_col = [&](int jj, I r0, I r1, I r2, I r3) {
                int4 o;
                o.x = (int32_t)conv_apply<I, true>(r0, outq);
                o.y = (int32_t)conv_apply<I, true>(r1, outq);
                o.z = (int32_t)conv_apply<I, true>(r2, outq);
                o.w = (int32_t)conv_apply<I, true>(r3, outq);
                if constexpr (Cfg::EPI) {                         // fused ANUS epilogue (qb200_gemul_poly_batched)
                    o.x = (int32_t)poly_ct<Cfg, I>((I)o.x); o.y = (int32_t)poly_ct<Cfg, I>((I)o.y);
                    o.z = (int32_t)poly_ct<Cfg, I>((I)o.z); o.w = (int32_t)poly_ct<Cfg, I>((I)o.w);
                }
                *(int4 *)(Ci + (j0 + jj) * M + i0) = o;
            };
            if constexpr (s_form_ok<Cfg>()) {
                if constexpr (T::SUBST) {
                    tile_compute_s<Cfg, false, QB_SFORM_JW>(Ag, Bg, i0, j0, one, store_col);
                } else if constexpr (DYN) {
                    if (clamp) tile_compute_s<Cfg, true, QB_SFORM_JW>(Ag, Bg, i0, j0, one, store_col);
                    else       tile_compute_s<Cfg, false, QB_SFORM_JW>(Ag, Bg, i0, j0, one, store_col);
                } else {
                    tile_compute_s<Cfg, true, QB_SFORM_JW>(Ag, Bg, i0, j0, one, store_col);
                }
            } else {
                I res[4][4];
                if constexpr (T::SUBST) {
                    tile_compute<Cfg, false>(Ag, Bg, i0, j0, res);
                } else if constexpr (DYN) {
                    if (clamp) tile_compute<Cfg, true>(Ag, Bg, i0, j0, res);
                    else       tile_compute<Cfg, false>(Ag, Bg, i0, j0, res);
                } else {
                    tile_compute<Cfg, true>(Ag, Bg, i0, j0, res);
                }
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) store_col(jj, res[0][jj], res[1][jj], res[2][jj], res[3][jj]);
            }
        }
        __syncthreads();      // everyone is done reading the operands (and range_smem) before they are overwritten
    }
}

template <class Cfg>
__global__ void __launch_bounds__(GemulTile<Cfg, 4, 4>::BLOCK, gemul_min_blocks(GemulTile<Cfg, 4, 4>::BLOCK, GemulTile<Cfg, 4, 4>::SMEM_BYTES))
k_gemul_tile(const typename Cfg::WA *__restrict__ A, const typename Cfg::WB *__restrict__ B, int32_t *__restrict__ C, uint64_t batch)
{
    gemul_tile_body<Cfg>(A, B, C, batch);
}

#ifndef QB_JIT
template <class Cfg>
static int launch_gemul_tile(const void *A, const void *B, void *C, uint64_t batch, cudaStream_t s)
{
    typedef GemulTile<Cfg, 4, 4> T;
    static int grid_caps[64] = { 0 };                 // per device: function attributes and occupancy belong to a context
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    int &grid_cap = grid_caps[cur_dev & 63];
    if (!grid_cap) {
        cudaError_t e = cudaFuncSetAttribute(k_gemul_tile<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gemul_tile<Cfg>, T::BLOCK, T::SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        grid_cap = sms * (per_sm > 0 ? per_sm : 1);
    }
    const uint64_t ngroups = (batch + T::G - 1) / T::G;
    const unsigned grid = (unsigned)(ngroups < (uint64_t)grid_cap ? ngroups : (uint64_t)grid_cap);
    k_gemul_tile<Cfg><<<grid, T::BLOCK, T::SMEM_BYTES, s>>>((const typename Cfg::WA *)A, (const typename Cfg::WB *)B, (int32_t *)C, batch);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) count_launch(1);
    return (int)e;
}

#endif  // QB_JIT

}  // namespace qb
