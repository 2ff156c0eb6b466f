// qb_plan.cpp — host plan compiler (no CUDA in this file).  See qb_plan.h.
#include "qb_plan.h"
#include <cmath>
#include <algorithm>
#include <cstring>
#include <vector>

namespace qb {

static i128 shl(i128 v, int s) { return v * ((i128)1 << s); }

static int bits_for(Ival r)
{
    int n = 1;
    while (n < 127 && !(r.lo >= -((i128)1 << (n - 1)) && r.hi <= ((i128)1 << (n - 1)) - 1)) n++;
    return n;
}
static Ival hull(Ival a, Ival b) { return { std::min(a.lo, b.lo), std::max(a.hi, b.hi) }; }

int width(qb_qdesc t) { return 1 + t.int_bits + t.frac_bits; }            // QuBLAS.h:2384-2385
int natural_word(qb_qdesc t) { return width(t) <= 32 ? 4 : 8; }            // QuBLAS.h:353
int word_bytes(qb_qdesc t) { return t.word_bytes ? t.word_bytes : natural_word(t); }
Ival type_range(qb_qdesc t)
{
    int W = width(t);
    return { -((i128)1 << (W - 1)), ((i128)1 << (W - 1)) - 1 };
}

bool check_type(qb_qdesc t, Analysis &an, const char *what)
{
    if (t.int_bits + t.frac_bits < 0) {                                    // static_assert QuBLAS.h:2372
        an.fail(QB_EINVAL, std::string(what) + ": intBits + fracBits must be non-negative"); return false;
    }
    if (width(t) > 64) {
        an.fail(QB_EDOMAIN, std::string(what) + ": types wider than 64 bits are outside the GPU parity domain"); return false;
    }
    if (t.qu_mode > 6) { an.fail(QB_EINVAL, std::string(what) + ": unknown QuMode id"); return false; }
    if (t.of_mode > 3) { an.fail(QB_EINVAL, std::string(what) + ": OfMode not implemented (WRP::TCPL_SAT is a stub upstream, QuBLAS.h:2336-2344)"); return false; }
    if (t.word_bytes != 0 && t.word_bytes != 1 && t.word_bytes != 2 && t.word_bytes != 4 && t.word_bytes != 8) {
        an.fail(QB_EINVAL, std::string(what) + ": word_bytes must be 0/1/2/4/8"); return false;
    }
    if (t.word_bytes && t.word_bytes * 8 < width(t)) {
        an.fail(QB_EINVAL, std::string(what) + ": word_bytes too small for the type's width"); return false;
    }
    return true;
}

// AddMerger without to-args, QuBLAS.h:3125-3140.
qb_qdesc add_merge_default(qb_qdesc a, qb_qdesc b)
{
    qb_qdesc r;
    r.int_bits = std::max(a.int_bits, b.int_bits);
    r.frac_bits = std::max(a.frac_bits, b.frac_bits);
    r.is_signed = a.is_signed || b.is_signed;
    r.qu_mode = a.qu_mode == b.qu_mode ? a.qu_mode : (uint8_t)TRN_TCPL;
    r.of_mode = a.of_mode == b.of_mode ? a.of_mode : (uint8_t)SAT_TCPL;
    r.word_bytes = 0;
    return r;
}

static i128 round_exact(i128 v, int s, int mode)                           // QuBLAS.h:2002-2204
{
    if (s <= 0) return shl(v, -s);
    i128 fl = v >> s, rem = v - shl(fl, s), half = (i128)1 << (s - 1);
    switch (mode) {
    case RND_POS_INF: return fl + (rem >= half);
    case RND_NEG_INF: return fl + (rem > half);
    case RND_ZERO:    return fl + (rem > half || (rem == half && v < 0));
    case RND_INF:     return fl + (rem > half || (rem == half && v > 0));
    case RND_CONV:    return rem > half ? fl + 1 : (rem < half ? fl : fl + (fl & 1));
    case TRN_SMGN:    return fl + (v < 0 && rem != 0);
    default:          return fl;
    }
}

QStage make_stage(int from_frac, Ival in, int in_width, qb_qdesc to, TVal *out, Analysis &an)
{
    QStage st;
    memset(&st, 0, sizeof(st));
    int s = from_frac - to.frac_bits;
    if (s > 100 || s < -62) { an.fail(QB_EDOMAIN, "fractional shift out of range"); s = 0; }
    st.shift = (int8_t)s; st.qu = to.qu_mode; st.of = to.of_mode;
    an.max_shift = std::max(an.max_shift, s < 0 ? -s : s);
    an.max_bits = std::max(an.max_bits, bits_for(in));
    an.ref_width = std::max(an.ref_width, in_width);
    int stage_bits = bits_for(in);
    // Two-word slice: the reference instantiates a 65..128-bit ArbiInt for this stage.  Its wide path equals the
    // ideal model (tests/golden/ref_wide.json) except for RND::CONV (SURVEY 2.3 defect 3) and for values that reach
    // +-[2^63, 2^64) at the overflow compare (defect 4); left shifts beyond 64 bits are untested and fenced too.
    const bool ref_wide = in_width > 64;
    Ival r;
    if (s <= 0) {
        r = { shl(in.lo, -s), shl(in.hi, -s) };
        an.max_bits = std::max(an.max_bits, bits_for(r));
        stage_bits = std::max(stage_bits, bits_for(r));
        an.ref_width = std::max(an.ref_width, in_width - s);
        if (in_width - s > 64) an.fail(QB_EDOMAIN, "a left shift would create an intermediate wider than 64 bits (outside the parity domain)");
    } else {
        bool rnd = to.qu_mode <= RND_CONV;
        // Reference defects: ArbiInt<32/64>::allOnes() sign-extends and
        // staticShiftLeft<31/63> of a 1-bit value overflows the narrow word.
        if (rnd && (s == 32 || s == 64))
            an.fail(QB_EDOMAIN, "RND right shift of exactly 32/64 bits hits a reference defect (outside the parity domain)");
        if (to.qu_mode == RND_CONV && (s == 31 || s == 63))
            an.fail(QB_EDOMAIN, "RND::CONV right shift of 31/63 bits hits a reference defect (outside the parity domain)");
        if (s >= 63) an.fail(QB_EDOMAIN, "right shift >= 63 not supported");
        if (ref_wide && to.qu_mode == RND_CONV)
            an.fail(QB_EDOMAIN, "RND::CONV on an intermediate wider than 64 bits hits a reference defect (outside the parity domain)");
        r = { round_exact(in.lo, s, to.qu_mode), round_exact(in.hi, s, to.qu_mode) };
        if (ref_wide && bits_for(r) > 64)
            an.fail(QB_EDOMAIN, "a >64-bit intermediate can reach +-[2^63, 2^64) at the overflow compare (reference defect, outside the parity domain)");
        i128 bias = rnd ? ((i128)1 << (s - 1)) : 0;                       // device computes (v + bias) >> s
        Ival biased = { in.lo, in.hi + bias };
        an.max_bits = std::max(an.max_bits, bits_for(biased));
        stage_bits = std::max(stage_bits, bits_for(biased));
        an.lossless = false;
    }
    // intConvert, QuBLAS.h:2227-2334
    int W = width(to);
    i128 mx = ((i128)1 << (W - 1)) - 1;
    i128 mn = to.is_signed ? -((i128)1 << (W - 1)) : 0;
    Ival o = r;
    switch (to.of_mode) {
    case SAT_TCPL: case SAT_SMGN: {
        i128 lo = to.of_mode == SAT_SMGN ? (to.is_signed ? -mx : 0) : mn;
        st.lo = (int64_t)lo; st.hi = (int64_t)mx;
        st.need_lo = r.lo < lo; st.need_hi = r.hi > mx;
        o = { std::min(std::max(r.lo, lo), mx), std::max(std::min(r.hi, mx), lo) };
        break;
    }
    case SAT_ZERO: {
        st.lo = (int64_t)mn; st.hi = (int64_t)mx;
        st.need_lo = r.lo < mn; st.need_hi = r.hi > mx;
        if (st.need_lo || st.need_hi) {
            Ival in_rng = { std::max(r.lo, mn), std::min(r.hi, mx) };
            o = in_rng.lo <= in_rng.hi ? hull(in_rng, { 0, 0 }) : Ival{ 0, 0 };
        }
        break;
    }
    default: {  // WRP_TCPL
        st.wrap_signed = to.is_signed;
        st.wrap_bits = (uint8_t)(to.is_signed ? W : W - 1);
        i128 lo = to.is_signed ? mn : 0;
        st.lo = (int64_t)lo; st.hi = (int64_t)mx;
        bool out_of_range = r.lo < lo || r.hi > mx;
        st.need_lo = st.need_hi = out_of_range;
        if (out_of_range) {
            o = { lo, mx };
            // Reference defect found by oracle/ref_harness_gen.py: the unsigned wrap mask
            // ArbiInt<32/64>::allOnes() is -1 and sign-extends, so the mask is a no-op.
            if (!to.is_signed && (W - 1 == 32 || W - 1 == 64))
                an.fail(QB_EDOMAIN, "unsigned WRP::TCPL with intBits+fracBits of exactly 32/64 hits a reference defect");
        }
        break;
    }
    }
    if (st.need_lo || st.need_hi) an.lossless = false;
    an.max_bits = std::max(an.max_bits, bits_for(o));
    an.max_reg_bits = std::max(an.max_reg_bits, bits_for(o));
    st.wide = stage_bits > 64 ? 2 : ((stage_bits > 32 || s > 30 || s < -30) ? 1 : 0);   // 2: evaluate in __int128
    if (st.wide == 2) an.max_reg_bits = std::max(an.max_reg_bits, 33);                   // the __int128 path lives in the 64-bit register class
    st.identity = (s <= 0 && !st.need_lo && !st.need_hi);
    if (out) { out->t = to; out->t.word_bytes = 0; out->r = o; }
    return st;
}

QStage make_mul(TVal a, TVal b, qb_qdesc to, TVal *out, Analysis &an)      // Qmul_s::mul, QuBLAS.h:3146-3171
{
    i128 c[4] = { a.r.lo * b.r.lo, a.r.lo * b.r.hi, a.r.hi * b.r.lo, a.r.hi * b.r.hi };
    Ival p = { *std::min_element(c, c + 4), *std::max_element(c, c + 4) };
    an.max_reg_bits = std::max(an.max_reg_bits, std::max(bits_for(a.r), bits_for(b.r)));
    return make_stage(a.frac() + b.frac(), p, width(a.t) + width(b.t), to, out, an);
}

AddStage make_add(TVal a, TVal b, qb_qdesc to, bool sub, TVal *out, Analysis &an)  // QuBLAS.h:3177-3235
{
    AddStage s;
    memset(&s, 0, sizeof(s));
    int mf = std::max(a.frac(), b.frac());
    int sa = mf - a.frac(), sb = mf - b.frac();
    if (sa > 62 || sb > 62) { an.fail(QB_EDOMAIN, "alignment shift out of range"); sa = sb = 0; }
    s.shl_a = (uint8_t)sa; s.shl_b = (uint8_t)sb; s.sub = sub;
    an.max_shift = std::max(an.max_shift, std::max(sa, sb));
    Ival x = { shl(a.r.lo, sa), shl(a.r.hi, sa) }, y = { shl(b.r.lo, sb), shl(b.r.hi, sb) };
    an.max_bits = std::max(an.max_bits, std::max(bits_for(x), bits_for(y)));
    an.max_reg_bits = std::max(an.max_reg_bits, std::max(bits_for(a.r), bits_for(b.r)));
    Ival e = sub ? Ival{ x.lo - y.hi, x.hi - y.lo } : Ival{ x.lo + y.lo, x.hi + y.hi };
    int w = std::max(width(a.t) + sa, width(b.t) + sb) + 1;
    if (sub && w > 64) an.fail(QB_EDOMAIN, "Qsub on an intermediate wider than 64 bits hits a reference defect (outside the parity domain)");
    s.q = make_stage(mf, e, w, to, out, an);
    if (s.q.wide == 0 && (std::max(bits_for(x), bits_for(y)) > 32 || sa > 30 || sb > 30)) s.q.wide = 1;
    return s;
}

static qb_qdesc level_type(const qb_levels &L, unsigned layer, qb_qdesc a, qb_qdesc b)
{
    if (L.n == 0) return add_merge_default(a, b);                          // QuBLAS.h:4917-4921
    return L.lv[layer < L.n ? layer : L.n - 1];                            // QuBLAS.h:4913
}

TVal build_tree(uint64_t count, TVal leaf, const qb_levels &L, int form, TreeProg *prog, Analysis &an)
{
    memset(prog, 0, sizeof(*prog));
    if (count == 0) { an.fail(QB_EINVAL, "empty reduction"); return leaf; }
    if (L.n > QB_MAX_LEVELS) { an.fail(QB_EINVAL, "too many levels"); return leaf; }
    for (uint32_t i = 0; i < L.n; i++) if (!check_type(L.lv[i], an, "level type")) return leaf;
    unsigned fl = 0; while ((count >> (fl + 1)) != 0) fl++;               // floor(log2(count))
    bool pow2 = (count & (count - 1)) == 0;
    unsigned cl = pow2 ? fl : fl + 1;                                      // ceil(log2(count))
    if (cl > QBP_MAX_LEVELS) { an.fail(QB_EDOMAIN, "reduction too deep"); return leaf; }
    std::vector<TVal> node(cl + 1);
    node[0] = leaf;
    prog->nlev = fl;
    for (unsigned l = 0; l < cl; l++) {
        qb_qdesc lt = level_type(L, l, node[l].t, node[l].t);
        TVal nx;
        prog->lvl[l] = make_add(node[l], node[l], lt, false, &nx, an);
        // an odd leftover is converted (not added) into the next level, QuBLAS.h:4977-4980; with no
        // Args the buffer has the element type itself (:4966) and the carry is a plain copy.
        if (L.n != 0) {
            TVal cv;
            Analysis scratch = an;
            prog->carry[l] = make_stage(node[l].frac(), node[l].r, width(node[l].t), lt, &cv, scratch);
            if (!pow2 && form == QB_TREE_VECTOR) { nx.r = hull(nx.r, cv.r); an = scratch; }
        } else {
            memset(&prog->carry[l], 0, sizeof(QStage));
            prog->carry[l].identity = 1; prog->carry[l].qu = TRN_TCPL;
            prog->carry[l].lo = (int64_t)type_range(lt).lo; prog->carry[l].hi = (int64_t)type_range(lt).hi;
            if (!pow2 && form == QB_TREE_VECTOR) nx.r = hull(nx.r, node[l].r);
        }
        node[l + 1] = nx;
    }
    if (pow2) { prog->top_slot = fl; prog->nflush = 0; return node[fl]; }

    // non-power-of-two: stack slots = set bits of count
    std::vector<unsigned> slots;
    for (unsigned l = 0; l <= fl; l++) if ((count >> l) & 1) slots.push_back(l);   // ascending
    uint32_t nf = 0;
    auto push = [&](FlushStep s) { if (nf < QBP_MAX_FLUSH) prog->flush[nf++] = s; else an.fail(QB_EDOMAIN, "flush program too long"); };
    FlushStep st;
    TVal cur;
    if (form == QB_TREE_VARIADIC) {
        // deepest first; each leftover is added with ITS level's type, QuBLAS.h:4944-4948
        memset(&st, 0, sizeof(st)); st.kind = FL_LOAD; st.slot = (uint8_t)slots.back(); push(st);
        cur = node[slots.back()];
        for (int i = (int)slots.size() - 2; i >= 0; i--) {
            unsigned le = slots[i];
            TVal e = node[le];
            qb_qdesc lt = level_type(L, le, cur.t, e.t);
            memset(&st, 0, sizeof(st)); st.kind = FL_ADD; st.slot = (uint8_t)le;
            st.a = make_add(e, cur, lt, false, &cur, an);
            push(st);
        }
    } else {
        // vector form: the leftover climbs, converted at every level it passes, QuBLAS.h:4977-4980
        memset(&st, 0, sizeof(st)); st.kind = FL_LOAD; st.slot = (uint8_t)slots[0]; push(st);
        cur = node[slots[0]];
        unsigned curl = slots[0];
        // the leaf-level entry itself has the exact leaf interval
        for (size_t i = 1; i < slots.size(); i++) {
            unsigned le = slots[i];
            while (curl < le) {
                if (L.n != 0) {
                    qb_qdesc lt = level_type(L, curl, cur.t, cur.t);
                    memset(&st, 0, sizeof(st)); st.kind = FL_CONV;
                    st.a.q = make_stage(cur.frac(), cur.r, width(cur.t), lt, &cur, an);
                    push(st);
                }
                curl++;
            }
            TVal e = node[le];
            qb_qdesc lt = level_type(L, le, e.t, cur.t);
            memset(&st, 0, sizeof(st)); st.kind = FL_ADD; st.slot = (uint8_t)le;
            st.a = make_add(e, cur, lt, false, &cur, an);
            push(st);
            curl = le + 1;
        }
    }
    prog->nflush = nf;
    prog->top_slot = 0;
    return cur;
}

TfProg build_tf(TVal a, TVal b, TVal c, TVal d, const qb_tfc_types &t, TVal *re, TVal *im, Analysis &an, TVal *mid)
{
    TfProg p;
    memset(&p, 0, sizeof(p));
    const qb_qdesc *all[8] = { &t.ab, &t.cd, &t.ba, &t.abc, &t.bad, &t.cdb, &t.AB, &t.BC };
    for (auto q : all) if (!check_type(*q, an, "TFComplexMul node type")) return p;
    TVal ab, cd, ba, A, B, C;
    p.ab = make_add(a, b, t.ab, false, &ab, an);          // QuBLAS.h:3524 Qadd<abT>(a, b)
    p.cd = make_add(c, d, t.cd, false, &cd, an);          // :3525 Qadd<cdT>(c, d)
    p.ba = make_add(b, a, t.ba, true, &ba, an);           // :3526 Qsub(b, a)
    if (mid) { mid[0] = ab; mid[1] = cd; mid[2] = ba; }
    p.A = make_mul(ab, c, t.abc, &A, an);                 // :3524
    p.B = make_mul(cd, b, t.bad, &B, an);                 // :3525
    p.C = make_mul(ba, d, t.cdb, &C, an);                 // :3526
    p.re = make_add(A, B, t.AB, true, re, an);            // :3528
    p.im = make_add(B, C, t.BC, true, im, an);            // :3529
    return p;
}

// 32-bit register class: bounds of wider target types are clamped to the int32 range (a no-op for
// values that provably fit 32 bits); the device casts them to int32.
static void sanitize32(QStage &q)
{
    if (q.lo < INT32_MIN) q.lo = INT32_MIN;
    if (q.hi > INT32_MAX) q.hi = INT32_MAX;
}
void sanitize32(TreeProg &t)
{
    for (int i = 0; i < QBP_MAX_LEVELS; i++) { sanitize32(t.lvl[i].q); sanitize32(t.carry[i]); }
    for (int i = 0; i < QBP_MAX_FLUSH; i++) sanitize32(t.flush[i].a.q);
}
void sanitize32(TfProg &t)
{
    sanitize32(t.ab.q); sanitize32(t.cd.q); sanitize32(t.ba.q); sanitize32(t.A); sanitize32(t.B); sanitize32(t.C);
    sanitize32(t.re.q); sanitize32(t.im.q);
}
void sanitize32_stage(QStage &q) { sanitize32(q); }

void sanitize32(BcmProg &t)
{
    sanitize32(t.ac); sanitize32(t.bd); sanitize32(t.ad); sanitize32(t.bc); sanitize32(t.re.q); sanitize32(t.im.q);
}

BcmProg build_bcm(TVal a, TVal b, TVal c, TVal d, const qb_bcm_types &t, TVal *re, TVal *im, Analysis &an)
{
    BcmProg p;
    memset(&p, 0, sizeof(p));
    const qb_qdesc *all[6] = { &t.ac, &t.bd, &t.ad, &t.bc, &t.acbd, &t.adbc };
    for (auto q : all) if (!check_type(*q, an, "BasicComplexMul node type")) return p;
    TVal ac, bd, ad, bc;
    p.ac = make_mul(a, c, t.ac, &ac, an);                 // QuBLAS.h:3439 Qmul<acT>(a, c)
    p.bd = make_mul(b, d, t.bd, &bd, an);                 //               Qmul<bdT>(b, d)
    p.ad = make_mul(a, d, t.ad, &ad, an);                 // :3440        Qmul<adT>(a, d)
    p.bc = make_mul(b, c, t.bc, &bc, an);                 //               Qmul<bcT>(b, c)
    p.re = make_add(ac, bd, t.acbd, true, re, an);        // :3439 Qsub<acbdT>
    p.im = make_add(ad, bc, t.adbc, false, im, an);       // :3440 Qadd<adbcT>
    return p;
}

// real (x) complex (QuBLAS.h:3600-3637): two independent products.  kind = QB_CPROD_REAL_A (x = a real,
// y = c + di) or QB_CPROD_REAL_B (x = a + bi, y = c real); unused stages stay zero.
BcmProg build_rc(int kind, TVal a, TVal b, TVal c, TVal d, const qb_bcm_types &t, TVal *re, TVal *im, Analysis &an)
{
    BcmProg p;
    memset(&p, 0, sizeof(p));
    const qb_qdesc &imt = kind == QB_CPROD_REAL_A ? t.ad : t.bc;
    if (!check_type(t.ac, an, "real-complex product type (real part)") || !check_type(imt, an, "real-complex product type (imaginary part)")) return p;
    p.ac = make_mul(a, c, t.ac, re, an);
    if (kind == QB_CPROD_REAL_A) p.ad = make_mul(a, d, t.ad, im, an);
    else p.bc = make_mul(b, c, t.bc, im, an);
    return p;
}

static void finish(Analysis &an)
{
    if (an.status) return;
    if (an.ref_width > 128)
        an.fail(QB_EDOMAIN, "an intermediate of the reference would be wider than 128 bits (outside the two-word slice)");
    else if (an.max_bits > 126)
        an.fail(QB_EDOMAIN, "an exact intermediate needs more than 126 bits");
}

int compile_reduce(const qb_reduce_plan &p, ReduceProg *out, Analysis &an)
{
    if (!check_type(p.elem, an, "element type")) return an.status;
    if (p.len == 0) { an.fail(QB_EINVAL, "len must be >= 1"); return an.status; }
    if (p.tree_form > 1) { an.fail(QB_EINVAL, "bad tree_form"); return an.status; }
    TVal leaf = { p.elem, type_range(p.elem) };
    leaf.t.word_bytes = 0;
    out->out = build_tree(p.len, leaf, p.levels, (int)p.tree_form, &out->tree, an);
    out->elem_wb = word_bytes(p.elem);
    out->out_wb = natural_word(out->out.t);
    finish(an);
    return an.status;
}

int compile_gemul(const qb_gemul_plan &p, GemulProg *out, Analysis &an)
{
    if (!check_type(p.a, an, "A type") || !check_type(p.b, an, "B type") || !check_type(p.mul, an, "MulArgs type") ||
        !check_type(p.c, an, "C type")) return an.status;
    if (!p.M || !p.N || !p.K) { an.fail(QB_EINVAL, "M, N, K must be >= 1"); return an.status; }
    if (p.tree_form > 1) { an.fail(QB_EINVAL, "bad tree_form"); return an.status; }
    TVal a = { p.a, type_range(p.a) }, b = { p.b, type_range(p.b) };
    out->mul = make_mul(a, b, p.mul, &out->prod, an);
    out->acc = build_tree(p.K, out->prod, p.add, p.tree_form, &out->tree, an);
    TVal c;
    out->out = make_stage(out->acc.frac(), out->acc.r, width(out->acc.t), p.c, &c, an);   // converting ctor :2398-2411
    out->a_wb = word_bytes(p.a); out->b_wb = word_bytes(p.b); out->c_wb = word_bytes(p.c);
    finish(an);
    return an.status;
}

static int cword(qb_qdesc re, qb_qdesc im) { return std::max(word_bytes(re), word_bytes(im)); }

int compile_cgemul(const qb_cgemul_plan &p, CGemulProg *out, Analysis &an)
{
    if (p.product > QB_CPROD_REAL_B) { an.fail(QB_EINVAL, "bad complex product kind"); return an.status; }
    const bool real_a = p.product == QB_CPROD_REAL_A, real_b = p.product == QB_CPROD_REAL_B;
    const qb_qdesc &a_im = real_a ? p.a_re : p.a_im, &b_im = real_b ? p.b_re : p.b_im;       // the ignored descriptor is not validated
    const qb_qdesc *ts[6] = { &p.a_re, &a_im, &p.b_re, &b_im, &p.c_re, &p.c_im };
    for (auto t : ts) if (!check_type(*t, an, "complex element type")) return an.status;
    if (!p.M || !p.N || !p.K) { an.fail(QB_EINVAL, "M, N, K must be >= 1"); return an.status; }
    if (p.tree_form > 1) { an.fail(QB_EINVAL, "bad tree_form"); return an.status; }
    TVal a = { p.a_re, type_range(p.a_re) }, b = { a_im, type_range(a_im) };
    TVal c = { p.b_re, type_range(p.b_re) }, d = { b_im, type_range(b_im) };
    TVal re, im;
    out->kind = p.product;
    memset(&out->tf, 0, sizeof(out->tf)); memset(&out->bcm, 0, sizeof(out->bcm));
    if (real_a || real_b) out->bcm = build_rc(p.product, a, b, c, d, p.bcm, &re, &im, an);
    else if (p.product == QB_CPROD_BASIC) out->bcm = build_bcm(a, b, c, d, p.bcm, &re, &im, an);
    else out->tf = build_tf(a, b, c, d, p.tf, &re, &im, an);
    if (an.status) return an.status;
    TVal ar = build_tree(p.K, re, p.add_re, p.tree_form, &out->tree_re, an);
    TVal ai = build_tree(p.K, im, p.add_im, p.tree_form, &out->tree_im, an);
    TVal o;
    out->out_re = make_stage(ar.frac(), ar.r, width(ar.t), p.c_re, &o, an);
    out->out_im = make_stage(ai.frac(), ai.r, width(ai.t), p.c_im, &o, an);
    out->a_wb = cword(p.a_re, a_im); out->b_wb = cword(p.b_re, b_im); out->c_wb = cword(p.c_re, p.c_im);
    finish(an);
    return an.status;
}

int compile_tfcmul(const qb_tfcmul_plan &p, TfProg *out, BcmProg *bout, TVal *re, TVal *im, Analysis &an)
{
    memset(out, 0, sizeof(*out)); memset(bout, 0, sizeof(*bout));
    if (p.product > QB_CPROD_REAL_B) { an.fail(QB_EINVAL, "bad complex product kind"); return an.status; }
    const bool real_x = p.product == QB_CPROD_REAL_A, real_y = p.product == QB_CPROD_REAL_B;
    const qb_qdesc &x_im = real_x ? p.x_re : p.x_im, &y_im = real_y ? p.y_re : p.y_im;
    const qb_qdesc *ts[4] = { &p.x_re, &x_im, &p.y_re, &y_im };
    for (auto t : ts) if (!check_type(*t, an, "complex element type")) return an.status;
    TVal a = { p.x_re, type_range(p.x_re) }, b = { x_im, type_range(x_im) };
    TVal c = { p.y_re, type_range(p.y_re) }, d = { y_im, type_range(y_im) };
    if (real_x || real_y) *bout = build_rc((int)p.product, a, b, c, d, p.bcm, re, im, an);
    else if (p.product == QB_CPROD_BASIC) *bout = build_bcm(a, b, c, d, p.bcm, re, im, an);
    else *out = build_tf(a, b, c, d, p.tf, re, im, an);
    finish(an);
    return an.status;
}

// Lossless predicate for the tcgen05 kind::i8 path: operand raws fit int8,
// every quantizer between the exact products and the tree result is an identity with shift 0,
// so the tree result IS the exact integer dot product (order-independent) and fits int32.
// ANUS::Qpoly<a0, a1, ...>(x) = Qadd<A0>(a0, Qmul<A0>(x, Qpoly<a1, ...>(x))) (QuBLAS.h:4846-4853), innermost first;
// Qapprox picks the first Segment whose breakpoint exceeds x.toDouble() and converts the value into x's type (:4866-4888).
int compile_poly(const qb_poly_plan &p, PolyProg *out, qb_qdesc *out_type, Analysis &an)
{
    memset(out, 0, sizeof(*out));
    if (!check_type(p.x, an, "x type")) return an.status;
    if (p.nseg < 1 || p.nseg > QB_POLY_MAX_SEG) { an.fail(QB_EINVAL, "1 .. 8 segments"); return an.status; }
    if (!p.approx && p.nseg != 1) { an.fail(QB_EINVAL, "Qpoly takes one segment; several segments need approx = 1 (Qapprox)"); return an.status; }
    if (p.approx && width(p.x) > 53) { an.fail(QB_EDOMAIN, "Qapprox compares x.toDouble(): x wider than 53 bits would round"); return an.status; }
    out->nseg = p.nseg; out->approx = p.approx ? 1 : 0;
    const TVal x = { p.x, type_range(p.x) };
    an.max_reg_bits = std::max(an.max_reg_bits, width(p.x));
    for (uint32_t s = 0; s < p.nseg; s++) {
        const qb_poly_segment &sg = p.seg[s];
        PolySeg &ps = out->seg[s];
        if (sg.ncoef < 1 || sg.ncoef > QB_POLY_MAX_COEF) { an.fail(QB_EINVAL, "1 .. 8 coefficients per segment"); return an.status; }
        ps.ncoef = sg.ncoef;
        // x.toDouble() < breakpoint  <=>  raw < ceil(breakpoint * 2^fracBits)   (toDouble is exact up to 53 bits)
        if (s + 1 < p.nseg) {
            const long double t = ceill(ldexpl((long double)sg.breakpoint, p.x.frac_bits));
            ps.below = sg.breakpoint != sg.breakpoint ? INT64_MIN : t >= 9.2e18L ? INT64_MAX : t <= -9.2e18L ? INT64_MIN : (int64_t)t;
        } else {
            ps.below = INT64_MAX;
        }
        for (uint32_t i = 0; i < sg.ncoef; i++) {
            if (!check_type(sg.coef[i].type, an, "coefficient type")) return an.status;
            const Ival r = type_range(sg.coef[i].type);
            if ((i128)sg.coef[i].raw < r.lo || (i128)sg.coef[i].raw > r.hi) { an.fail(QB_EINVAL, "coefficient raw value outside its type"); return an.status; }
        }
        const qb_poly_coef &top = sg.coef[sg.ncoef - 1];
        TVal acc = { top.type, { (i128)top.raw, (i128)top.raw } };
        acc.t.word_bytes = 0;
        ps.top = top.raw;
        an.max_reg_bits = std::max(an.max_reg_bits, bits_for(acc.r));
        for (int i = (int)sg.ncoef - 2; i >= 0; i--) {
            const qb_poly_coef &c = sg.coef[i];
            TVal prod, ai = { c.type, { (i128)c.raw, (i128)c.raw } };
            ai.t.word_bytes = 0;
            ps.step[i].mul = make_mul(x, acc, c.type, &prod, an);
            ps.step[i].add = make_add(ai, prod, c.type, false, &acc, an);
            ps.step[i].a = c.raw;
        }
        if (p.approx) {
            TVal o;
            ps.back = make_stage(acc.frac(), acc.r, width(acc.t), p.x, &o, an);       // decltype(x){ ... }: converting ctor :2398-2411
        }
        if (an.status) return an.status;
    }
    if (out_type) { *out_type = p.approx ? p.x : p.seg[0].coef[0].type; out_type->word_bytes = p.approx ? p.x.word_bytes : p.seg[0].coef[0].type.word_bytes; }
    finish(an);
    return an.status;
}
void sanitize32(PolyProg &t)
{
    for (auto &sg : t.seg) { for (auto &st : sg.step) { sanitize32(st.mul); sanitize32(st.add.q); } sanitize32(sg.back); }
}

bool gemul_tensor_core_ok(const qb_gemul_plan &p, const GemulProg &g, const Analysis &an)
{
    if (an.status) return false;
    if (width(p.a) > 8 || width(p.b) > 8) return false;
    if (!g.mul.identity || g.mul.shift != 0) return false;
    unsigned cl = 0; while (((uint64_t)1 << cl) < p.K) cl++;
    for (unsigned l = 0; l < cl; l++) {
        const AddStage &s = g.tree.lvl[l];
        if (!s.q.identity || s.q.shift != 0 || s.shl_a || s.shl_b) return false;
    }
    for (uint32_t i = 0; i < g.tree.nflush; i++) {
        const FlushStep &f = g.tree.flush[i];
        if (f.kind == FL_LOAD) continue;
        if (!f.a.q.identity || f.a.q.shift != 0 || f.a.shl_a || f.a.shl_b) return false;
    }
    i128 bound = (i128)p.K * 128 * 128;
    return bound <= (((i128)1 << 31) - 1);
}

}  // namespace qb
