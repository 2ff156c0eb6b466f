// qb_device.cuh — register-level quantizer primitives (integer shift / mask / clamp only).
//
// Every function is bit-exact with the reference's fracConvert (QuBLAS.h:2002-2204) followed by
// intConvert (QuBLAS.h:2227-2334) inside the parity domain checked by the host plan compiler.
// `CT` = the program is a compile-time constant (specialised kernels): need_* flags and modes fold
// away.  With CT=false the same code reads the program from the kernel's constant bank.
#pragma once
#include <stdint.h>
#include "qb_prog.h"

#define QB_DEV __device__ __forceinline__

namespace qb {

template <typename I> struct ITraits;
template <> struct ITraits<int32_t> { typedef uint32_t U; static constexpr int BITS = 32; };
template <> struct ITraits<int64_t> { typedef uint64_t U; static constexpr int BITS = 64; };
#ifndef QB_JIT
template <> struct ITraits<__int128> { typedef unsigned __int128 U; static constexpr int BITS = 128; };   // two-word intermediates
#endif

template <typename I> QB_DEV I imin(I a, I b) { return a < b ? a : b; }
template <typename I> QB_DEV I imax(I a, I b) { return a > b ? a : b; }

template <typename I> QB_DEV I shl_exact(I v, int s)
{
    typedef typename ITraits<I>::U U;
    return (I)((U)v << s);
}

// fracConvert: s = fromFrac - toFrac.  All modes are (v + bias) >> s with
//   TRN::TCPL bias 0 | TRN::SMGN (2^s - 1) for v < 0 | RND::POS_INF half | RND::NEG_INF half-1
//   RND::ZERO half-1 + (v<0) | RND::INF half - (v<0) | RND::CONV half-1 + lsb(v >> s)
template <typename I> QB_DEV I q_round(I v, int s, int mode)
{
    typedef typename ITraits<I>::U U;
    constexpr int B = ITraits<I>::BITS;
    if (s <= 0) return shl_exact<I>(v, -s);
    const I half = (I)((U)1 << (s - 1));
    switch (mode) {
    case TRN_TCPL:    return v >> s;
    case TRN_SMGN:    return (v + ((v >> (B - 1)) & (I)(((U)1 << s) - 1))) >> s;
    case RND_POS_INF: return (v + half) >> s;
    case RND_NEG_INF: return (v + (half - 1)) >> s;
    case RND_ZERO:    return (v + (half - 1) + (I)((U)v >> (B - 1))) >> s;
    case RND_INF:     return (v + half - (I)((U)v >> (B - 1))) >> s;
    default:                                                              // RND_CONV
        if (s == 1) { const I q = v >> 1; return q + (v & q & 1); }       // halves tie to the even neighbour: one op fewer
        return (v + (half - 1) + ((v >> s) & 1)) >> s;
    }
}

// intConvert.  Bounds and wrap width come from the program; clamps proven dead by the host's
// interval analysis are skipped when the program is a compile-time constant.
template <typename I, bool CT> QB_DEV I q_ovf(I v, const QStage &s)
{
    typedef typename ITraits<I>::U U;
    constexpr int B = ITraits<I>::BITS;
    const bool nlo = CT ? (s.need_lo != 0) : true, nhi = CT ? (s.need_hi != 0) : true;
    switch (s.of) {
    case SAT_TCPL: case SAT_SMGN:
        if (nhi) v = imin<I>(v, (I)s.hi);
        if (nlo) v = imax<I>(v, (I)s.lo);
        return v;
    case SAT_ZERO: {
        bool bad = (nhi && v > (I)s.hi) || (nlo && v < (I)s.lo);
        return bad ? (I)0 : v;
    }
    default:  // WRP_TCPL: a wrap that interval analysis proved unnecessary is skipped in both modes
              // (the target may be wider than the register class)
        if (!(s.need_lo || s.need_hi)) return v;
        if (s.wrap_signed) { const int k = B - s.wrap_bits; return (I)(shl_exact<I>(v, k)) >> k; }
        return v & (I)s.hi;
    }
}

template <typename I, bool CT> QB_DEV I q_apply(I v, const QStage &s)
{
    return q_ovf<I, CT>(q_round<I>(v, s.shift, s.qu), s);
}

// A value held in a register of the kernel's class I, converted by stage s (converting ctor).
// `wide` stages (exact input wider than 32 bits) are evaluated in 64 bits inside a 32-bit kernel.
template <typename I, bool CT> QB_DEV I conv_apply(I v, const QStage &s)
{
    if constexpr (sizeof(I) == 4) { if (s.wide) return (I)q_apply<int64_t, CT>((int64_t)v, s); }
#ifndef QB_JIT
    if constexpr (sizeof(I) == 8) { if (s.wide == 2) return (I)q_apply<__int128, CT>((__int128)v, s); }
#endif
    return q_apply<I, CT>(v, s);
}

// Qmul_s::mul, QuBLAS.h:3146-3171: exact product, then the stage.
template <typename I, bool CT> QB_DEV I mul_apply(I a, I b, const QStage &s)
{
    if constexpr (sizeof(I) == 4) { if (s.wide) return (I)q_apply<int64_t, CT>((int64_t)a * (int64_t)b, s); }
#ifndef QB_JIT
    if constexpr (sizeof(I) == 8) { if (s.wide == 2) return (I)q_apply<__int128, CT>((__int128)a * (__int128)b, s); }   // 65..126-bit product
#endif
    return q_apply<I, CT>(a * b, s);
}

// Qadd_s::add / Qsub_s::sub, QuBLAS.h:3177-3235: align, exact sum, then the stage.
template <typename I, bool CT> QB_DEV I add_apply(I a, I b, const AddStage &s)
{
    if constexpr (sizeof(I) == 4) {
        if (s.q.wide) {
            int64_t x = shl_exact<int64_t>((int64_t)a, s.shl_a), y = shl_exact<int64_t>((int64_t)b, s.shl_b);
            return (I)q_apply<int64_t, CT>(s.sub ? x - y : x + y, s.q);
        }
    }
#ifndef QB_JIT
    if constexpr (sizeof(I) == 8) {
        if (s.q.wide == 2) {
            __int128 x = shl_exact<__int128>((__int128)a, s.shl_a), y = shl_exact<__int128>((__int128)b, s.shl_b);
            return (I)q_apply<__int128, CT>(s.sub ? x - y : x + y, s.q);
        }
    }
#endif
    I x = shl_exact<I>(a, s.shl_a), y = shl_exact<I>(b, s.shl_b);
    return q_apply<I, CT>(s.sub ? x - y : x + y, s.q);
}

// Same-type tree merge (both operands are level-l nodes): no alignment shifts.
template <typename I, bool CT> QB_DEV I merge_apply(I a, I b, const AddStage &s)
{
    if constexpr (sizeof(I) == 4) { if (s.q.wide) return (I)q_apply<int64_t, CT>((int64_t)a + (int64_t)b, s.q); }
#ifndef QB_JIT
    if constexpr (sizeof(I) == 8) { if (s.q.wide == 2) return (I)q_apply<__int128, CT>((__int128)a + (__int128)b, s.q); }
#endif
    return q_apply<I, CT>(a + b, s.q);
}

// TFComplexMul::mul, QuBLAS.h:3522-3534.
template <typename I, bool CT>
QB_DEV void tf_apply(I a, I b, I c, I d, const TfProg &t, I &re, I &im)
{
    I ab = add_apply<I, CT>(a, b, t.ab);
    I cd = add_apply<I, CT>(c, d, t.cd);
    I ba = add_apply<I, CT>(b, a, t.ba);
    I A = mul_apply<I, CT>(ab, c, t.A);
    I Bv = mul_apply<I, CT>(cd, b, t.B);
    I Cv = mul_apply<I, CT>(ba, d, t.C);
    re = add_apply<I, CT>(A, Bv, t.re);
    im = add_apply<I, CT>(Bv, Cv, t.im);
}

// BasicComplexMul::mul, QuBLAS.h:3437-3445.
template <typename I, bool CT>
QB_DEV void bcm_apply(I a, I b, I c, I d, const BcmProg &t, I &re, I &im)
{
    I ac = mul_apply<I, CT>(a, c, t.ac), bd = mul_apply<I, CT>(b, d, t.bd);
    I ad = mul_apply<I, CT>(a, d, t.ad), bc = mul_apply<I, CT>(b, c, t.bc);
    re = add_apply<I, CT>(ac, bd, t.re);
    im = add_apply<I, CT>(ad, bc, t.im);
}

// One complex-valued product of the kind the plan names (enum qb_cproduct); for the real (x) complex
// kinds (QuBLAS.h:3600-3637) the absent imaginary operand is never read.
template <typename I, bool CT>
QB_DEV void cprod_apply(int kind, I a, I b, I c, I d, const TfProg &tf, const BcmProg &bcm, I &re, I &im)
{
    if (kind == 0) tf_apply<I, CT>(a, b, c, d, tf, re, im);
    else if (kind == 1) bcm_apply<I, CT>(a, b, c, d, bcm, re, im);
    else {
        re = mul_apply<I, CT>(a, c, bcm.ac);
        im = kind == 2 ? mul_apply<I, CT>(a, d, bcm.ad) : mul_apply<I, CT>(b, c, bcm.bc);
    }
}

// ANUS::Qpoly / Qapprox on one element (QuBLAS.h:4846-4853, :4866-4888): segment selection on the raw value, Horner chain
// with both quantizers of every step, optional conversion back into x's type.
template <typename I, bool CT>
QB_DEV I poly_apply(const PolyProg &P, I x)
{
    uint32_t s = 0;
    while (s + 1 < P.nseg && !((int64_t)x < P.seg[s].below)) ++s;          // first segment whose breakpoint exceeds x, else the last
    const PolySeg &S = P.seg[s];
    I acc = (I)S.top;
    for (int i = (int)S.ncoef - 2; i >= 0; --i) {
        const I m = mul_apply<I, CT>(x, acc, S.step[i].mul);
        acc = add_apply<I, CT>((I)S.step[i].a, m, S.step[i].add);
    }
    if (P.approx) acc = conv_apply<I, CT>(acc, S.back);
    return acc;
}

// The same with the program as a compile-time constant (Cfg::poly()): segments are tested one after the other, every
// Horner step is unrolled, all quantizers are immediates.
template <class Cfg, typename I, int S>
QB_DEV I poly_seg_ct(I x)
{
    constexpr PolySeg sg = Cfg::poly().seg[S];
    I acc = (I)sg.top;
#pragma unroll
    for (int i = (int)sg.ncoef - 2; i >= 0; --i) {
        const I m = mul_apply<I, true>(x, acc, sg.step[i].mul);
        acc = add_apply<I, true>((I)sg.step[i].a, m, sg.step[i].add);
    }
    if constexpr (Cfg::poly().approx != 0) acc = conv_apply<I, true>(acc, sg.back);
    return acc;
}
template <class Cfg, typename I, int S = 0>
QB_DEV I poly_ct(I x)
{
    if constexpr (S + 1 >= (int)Cfg::poly().nseg) return poly_seg_ct<Cfg, I, S>(x);
    else return (int64_t)x < Cfg::poly().seg[S].below ? poly_seg_ct<Cfg, I, S>(x) : poly_ct<Cfg, I, S + 1>(x);
}

// the polynomial fused into the LOADS of a Qgemul operand (Cfg::PRO_A / PRO_B: ANUS::Qpoly / Qapprox as a prologue)
template <class Cfg, int WHICH> struct PolySel {
    static constexpr __host__ __device__ PolyProg poly() { if constexpr (WHICH == 1) return Cfg::poly_a(); else return Cfg::poly_b(); }
};

// Qmul through the FP32 FMA pipe (see qb_codegen.cpp:fp_mul_ok): fma(a, b, 1.5 * 2^(23+S)) is the exact product
// plus ONE IEEE rounding at ulp 2^S, i.e. the rounded right shift by S (round-to-nearest-even == RND::CONV,
// round toward -inf == TRN::TCPL).  Returns the quantized product BIASED by the constant's bit pattern
// fp_kbits<S>(); clamps are applied on the biased bits (monotone).
template <int S> __host__ __device__ constexpr int fp_kbits() { return ((150 + S) << 23) | 0x400000; }   // bits of 1.5 * 2^(23+S)
template <int S, bool CT> QB_DEV int fp_mul_apply(float a, float b, const QStage &q)
{
    constexpr float Cm = (float)(3ull << (22 + S));                    // 1.5 * 2^(23+S), exact
    const float t = q.qu == RND_CONV ? __fmaf_rn(a, b, Cm) : __fmaf_rd(a, b, Cm);
    int bits = __float_as_int(t);
    if (q.need_hi) bits = imin<int>(bits, (int)q.hi + fp_kbits<S>());
    if (q.need_lo) bits = imax<int>(bits, (int)q.lo + fp_kbits<S>());
    return bits;
}

// Saturation by operand substitution (round 2).  When the ONLY product that Qmul's SAT clamp can touch is min x min of two
// full-range signed operands (a = -2^pa, b = -2^pb, RND::CONV, shift S <= pa, pb; the product quantizes to hi + 1 exactly),
// replacing every operand that equals its type's minimum by
//     a' = -(2^pa - 2^(S-pb-1)),   b' = -(2^pb - 2^(S-pa-1))
// in the float copy makes the plain FMA deliver the CLAMPED result, and no clamp instruction is left:
//   * a' * b' / 2^S = hi + 1 - 1 + 2^(S-pa-pb-2): rounds to hi, the saturated value;
//   * a' * y / 2^S for any other y (|y| <= 2^pb - 1) = -2^(pa-S) y + y / 2^(pb+1): an integer plus less than half, so it
//     rounds to -2^(pa-S) y, the exact (unclamped, untouched) product quantized; likewise x * b';
//   * a' and b' are exact floats (pa + pb - S + 1 <= 24 significant bits).
// The FMA sees exact real arithmetic and rounds once, so these three statements are the whole proof; qb200_selftest_fp
// additionally sweeps every operand pair on the GPU (mode 2).  Config 3: pa = pb = 15, S = 8, a' = b' = -32767.99609375.
struct FpSubst { bool ok; float a_sub, b_sub; };
__host__ __device__ constexpr int qb_log2_exact(long long v)
{
    if (v <= 0) return -1;
    int p = 0;
    while ((v & 1) == 0) { v >>= 1; ++p; }
    return v == 1 ? p : -1;
}
__host__ __device__ constexpr FpSubst fp_subst(bool fp_mul, int S, const QStage &m, long long a_lo, long long a_hi, long long b_lo, long long b_hi)
{
    FpSubst r = { false, 0.f, 0.f };
#ifdef QB_NO_SUBST                                             // experiments: keep the per-instance choice between a clamped and a clamp-free loop
    return r;
#endif
    if (!fp_mul || m.qu != RND_CONV || !m.need_hi || m.need_lo || !(m.of == SAT_TCPL || m.of == SAT_SMGN)) return r;
    const int pa = qb_log2_exact(-a_lo), pb = qb_log2_exact(-b_lo);
    if (pa < 0 || pb < 0 || a_hi != -a_lo - 1 || b_hi != -b_lo - 1 || S < 1 || pa < S || pb < S || pa + pb - S + 1 > 24) return r;
    if (m.hi != (1LL << (pa + pb - S)) - 1) return r;
    // no other product reaches past hi: the next largest ones are a_lo * (b_lo + 1) and (a_lo + 1) * b_lo
    const long long p1 = a_lo * (b_lo + 1), p2 = (a_lo + 1) * b_lo, half = 1LL << (S - 1);
    if (((p1 + half) >> S) > m.hi || ((p2 + half) >> S) > m.hi) return r;
    r.ok = true;
    r.a_sub = -((float)(1LL << pa) - 1.0f / (float)(1LL << (pb + 1 - S)));
    r.b_sub = -((float)(1LL << pb) - 1.0f / (float)(1LL << (pa + 1 - S)));
    return r;
}

// Non-power-of-two counts: run the host-built flush program over the pending stack entries
// (vector form QuBLAS.h:4977-4980, variadic form :4942-4949).
template <typename I, bool CT>
QB_DEV I flush_apply(const TreeProg &P, const I *st)
{
    if (P.nflush == 0) return st[P.top_slot];
    I cur = 0;
    for (uint32_t i = 0; i < P.nflush; ++i) {
        const FlushStep &f = P.flush[i];
        if (f.kind == FL_LOAD) cur = st[f.slot];
        else if (f.kind == FL_CONV) cur = conv_apply<I, CT>(cur, f.a.q);
        else cur = add_apply<I, CT>(st[f.slot], cur, f.a);
    }
    return cur;
}

// Streaming adjacent-pair tree with a level-indexed stack: after leaf k, bit l of (k+1) set <=> st[l]
// holds a pending level-l node.  Reproduces the reference's pairing order exactly.
template <typename I, bool CT, int MAXL, class LeafFn>
QB_DEV I tree_stream(const TreeProg &P, uint64_t count, LeafFn leaf)
{
    I st[MAXL];
    for (uint64_t k = 0; k < count; ++k) {
        I v = leaf(k);
        unsigned l = 0;
        uint64_t m = k;
        while (m & 1) { v = merge_apply<I, CT>(st[l], v, P.lvl[l]); ++l; m >>= 1; }
        st[l] = v;
    }
    return flush_apply<I, CT>(P, st);
}

// Sign-extending loads of raw words (ArbiInt::data) of 1/2/4/8 bytes.
template <typename I> QB_DEV I load_word(const void *p, uint64_t idx, int wb)
{
    switch (wb) {
    case 1: return (I)((const int8_t *)p)[idx];
    case 2: return (I)((const int16_t *)p)[idx];
    case 4: return (I)((const int32_t *)p)[idx];
    default: return (I)((const int64_t *)p)[idx];
    }
}
template <typename I> QB_DEV void store_word(void *p, uint64_t idx, int wb, I v)
{
    switch (wb) {
    case 1: ((int8_t *)p)[idx] = (int8_t)v; break;
    case 2: ((int16_t *)p)[idx] = (int16_t)v; break;
    case 4: ((int32_t *)p)[idx] = (int32_t)v; break;
    default: ((int64_t *)p)[idx] = (int64_t)v; break;
    }
}

// Counter-based synthetic raw words; identical to oracle/qb_oracle.c:qbo_synth_word.
__host__ __device__ inline uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ inline int64_t synth_word(uint64_t key, uint64_t index, int64_t lo, uint64_t span)
{
    uint64_t h = mix64(key ^ index);
    if (span == 0) return (int64_t)h;
    return lo + (int64_t)__umul64hi(h, span);
}

}  // namespace qb
