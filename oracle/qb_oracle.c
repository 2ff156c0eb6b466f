/*
 * qb_oracle.c — CPU restatement of the QuBLAS fixed-point hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; it is the checker the CUDA path is compared against, never the product.
 *
 * Parity status: PINNED.  This restatement is checked (tests/test_oracle_*.py) against
 *   (1) the reference's own known-answer tests test/RND/ and test/TRN/ (40 KATs),
 *   (2) golden vectors generated HERE by compiling the unmodified reference header
 *       (/root/reference/include/QuBLAS.h) with g++ -std=c++23 — generator: oracle/ref_harness_gen.py,
 *       fixtures: tests/golden/ (json) — covering Qmul/Qadd/Qsub (all 7 QuModes x 4 OfModes),
 *       the converting ctor, Qreduce (vector and variadic form), TFComplexMul and the composed Qgemul.
 * `Qgemul` itself exists only in the reference's readme (readme.md:80-87); its semantics are the
 * composition Qmul -> Qreduce -> converting assignment, pinned through those primitives.
 *
 * Everything is exact integer arithmetic on __int128; the domain is "every intermediate fits
 * in 64 bits" (the reference equals the ideal model there).
 *
 * Each function cites the reference lines (include/QuBLAS.h) it restates.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/qb200.h"

typedef __int128 i128;

typedef struct { i128 raw; qb_qdesc t; } qval;          /* a typed fixed-point scalar */
typedef struct { qval re, im; } cval;                    /* Qcomplex, QuBLAS.h:2501-2617 */

static inline i128 shl128(i128 v, int s) { return v * ((i128)1 << s); }

/* fracConvert<from,to,QuMode>::convert, QuBLAS.h:2002-2204; s = fromFrac - toFrac. */
i128 qbo_round(i128 v, int s, int mode)
{
    if (s <= 0) return shl128(v, -s);                      /* :2011-2014 exact left shift */
    if (s > 120) s = 120;
    i128 fl = v >> s;                                      /* staticShiftRight = floor */
    i128 rem = v - shl128(fl, s);                          /* Xl = val & mask */
    i128 half = (i128)1 << (s - 1);                        /* T */
    switch (mode) {
    case QB_RND_POS_INF: return fl + (rem >= half);                          /* :2026 */
    case QB_RND_NEG_INF: return fl + (rem > half);                           /* :2056 */
    case QB_RND_ZERO:    return fl + (rem > half || (rem == half && v < 0)); /* :2086 */
    case QB_RND_INF:     return fl + (rem > half || (rem == half && v > 0)); /* :2116 */
    case QB_RND_CONV:                                                        /* :2137-2156 */
        if (rem > half) return fl + 1;
        if (rem < half) return fl;
        return fl + (fl & 1);
    case QB_TRN_TCPL:    return fl;                                          /* :2166 */
    case QB_TRN_SMGN:    return fl + (v < 0 && rem != 0);                    /* :2189-2199 */
    default:             return fl;
    }
}

/* intConvert<toInt,toFrac,toIsSigned,OfMode>::convert, QuBLAS.h:2227-2334. */
i128 qbo_ovf(i128 v, qb_qdesc t)
{
    int W = 1 + t.int_bits + t.frac_bits;                  /* sign bit always stored, :2384-2385 */
    i128 mx = ((i128)1 << (W - 1)) - 1;
    i128 mn = t.is_signed ? -((i128)1 << (W - 1)) : 0;
    switch (t.of_mode) {
    case QB_SAT_TCPL: return v > mx ? mx : (v < mn ? mn : v);                /* :2239-2250 */
    case QB_SAT_ZERO: return (v > mx || v < mn) ? 0 : v;                     /* :2265-2276 */
    case QB_SAT_SMGN: {                                                      /* :2287 min+1 */
        i128 lo = t.is_signed ? -mx : 0;
        return v > mx ? mx : (v < lo ? lo : v);
    }
    case QB_WRP_TCPL:
        if (t.is_signed) {                                                   /* :2312-2325 */
            i128 m = ((i128)1 << W) - 1;
            i128 low = v & m;
            if (low >> (W - 1)) low |= ~m;
            return low;
        }
        return v & mx;                                                       /* :2329-2331 */
    default: return v;
    }
}

/* Converting ctor: value `raw` with `from_frac` fractional bits -> type t, using t's modes
 * (QuBLAS.h:2398-2411). */
i128 qbo_conv(i128 raw, int from_frac, qb_qdesc t)
{
    return qbo_ovf(qbo_round(raw, from_frac - t.frac_bits, t.qu_mode), t);
}

static int imax(int a, int b) { return a > b ? a : b; }

/* AddMerger with no to-args (QuBLAS.h:3125-3140): int = max, frac = max, signed = or,
 * modes = common mode else TRN::TCPL / SAT::TCPL. */
qb_qdesc qbo_add_merge_default(qb_qdesc a, qb_qdesc b)
{
    qb_qdesc r;
    r.int_bits = (int16_t)imax(a.int_bits, b.int_bits);
    r.frac_bits = (int16_t)imax(a.frac_bits, b.frac_bits);
    r.is_signed = (uint8_t)(a.is_signed || b.is_signed);
    r.qu_mode = a.qu_mode == b.qu_mode ? a.qu_mode : QB_TRN_TCPL;
    r.of_mode = a.of_mode == b.of_mode ? a.of_mode : QB_SAT_TCPL;
    r.word_bytes = 0;
    return r;
}

/* Qmul_s::mul, QuBLAS.h:3146-3171 (to = resolved MulMerger type). */
qval qbo_qmul(qval a, qval b, qb_qdesc to)
{
    qval r; r.t = to;
    r.raw = qbo_conv(a.raw * b.raw, a.t.frac_bits + b.t.frac_bits, to);      /* :3157-3161 */
    return r;
}

/* Qadd_s::add / Qsub_s::sub, QuBLAS.h:3177-3235. */
static qval addsub(qval a, qval b, qb_qdesc to, int sub)
{
    int mf = imax(a.t.frac_bits, b.t.frac_bits);
    i128 x = shl128(a.raw, mf - a.t.frac_bits);                              /* shiftA :3182 */
    i128 y = shl128(b.raw, mf - b.t.frac_bits);                              /* shiftB :3183 */
    qval r; r.t = to;
    r.raw = qbo_conv(sub ? x - y : x + y, mf, to);                           /* :3190-3194 */
    return r;
}
qval qbo_qadd(qval a, qval b, qb_qdesc to) { return addsub(a, b, to, 0); }
qval qbo_qsub(qval a, qval b, qb_qdesc to) { return addsub(a, b, to, 1); }

/* Level type: TypeAt<min(layer, n-1)> (QuBLAS.h:4913); n == 0 -> default merger (:4917-4921). */
static qb_qdesc level_type(const qb_levels *L, unsigned layer, qb_qdesc a, qb_qdesc b)
{
    if (L->n == 0) return qbo_add_merge_default(a, b);
    return L->lv[layer < L->n ? layer : L->n - 1];
}

/* Vector form, QuBLAS.h:4960-4990: pairs (2i,2i+1) added with the level type; an odd leftover
 * is ASSIGNED (converted with the level type's modes) into the next level's buffer (:4977-4980;
 * the buffer has the element type itself when there are no Args, :4966).  All nodes of one level
 * share one type, so a level is an int64 array plus its type. */
static qval tree_vec(int64_t *xs, size_t n, qb_qdesc t, const qb_levels *L,
                     int64_t *partials, unsigned partial_level)
{
    unsigned layer = 0;
    if (partials && partial_level == 0) memcpy(partials, xs, n * sizeof(int64_t));
    while (n > 1) {
        qb_qdesc lt = level_type(L, layer, t, t);
        size_t h = n / 2;
        for (size_t i = 0; i < h; i++) {
            qval a = { xs[2 * i], t }, b = { xs[2 * i + 1], t };
            xs[i] = (int64_t)qbo_qadd(a, b, lt).raw;
        }
        if (n & 1) {
            xs[h] = L->n == 0 ? xs[n - 1] : (int64_t)qbo_conv(xs[n - 1], t.frac_bits, lt);
            h++;
        }
        n = h; layer++; t = lt;
        if (partials && layer == partial_level) memcpy(partials, xs, n * sizeof(int64_t));
    }
    qval r = { xs[0], t };
    return r;
}

/* Variadic form, QuBLAS.h:4924-4957: reduce the first 2*floor(n/2) elements recursively, THEN add the
 * leftover with THIS level's type (:4944-4948). */
static qval tree_var(int64_t *xs, size_t n, qb_qdesc t, const qb_levels *L, unsigned layer)
{
    qval last = { xs[n - 1], t };
    if (n == 1) return last;                                                 /* :4929-4932 */
    qb_qdesc lt = level_type(L, layer, t, t);
    size_t h = n / 2;
    for (size_t i = 0; i < h; i++) {
        qval a = { xs[2 * i], t }, b = { xs[2 * i + 1], t };
        xs[i] = (int64_t)qbo_qadd(a, b, lt).raw;
    }
    qval r = tree_var(xs, h, lt, L, layer + 1);
    if (n & 1) r = qbo_qadd(r, last, level_type(L, layer, r.t, last.t));
    return r;
}

static qval tree(int64_t *xs, size_t n, qb_qdesc t, const qb_levels *L, int form)
{
    return form == QB_TREE_VARIADIC ? tree_var(xs, n, t, L, 0) : tree_vec(xs, n, t, L, NULL, 0);
}

/* Result type of a tree over n elements of type t (no data needed). */
static qb_qdesc tree_type(size_t n, qb_qdesc t, const qb_levels *L, int form, unsigned layer)
{
    if (form == QB_TREE_VARIADIC) {
        if (n == 1) return t;
        qb_qdesc lt = level_type(L, layer, t, t);
        qb_qdesc r = tree_type(n / 2, lt, L, form, layer + 1);
        if (n & 1) r = level_type(L, layer, r, t);
        return r;
    }
    while (n > 1) { t = level_type(L, layer, t, t); n = (n + 1) / 2; layer++; }
    return t;
}

/* ---- raw buffer helpers: words are the reference's ArbiInt::data (QuBLAS.h:353) ---- */
static int word_bytes_of(qb_qdesc t)
{
    if (t.word_bytes) return t.word_bytes;
    return (1 + t.int_bits + t.frac_bits) <= 32 ? 4 : 8;
}
static i128 load_word(const void *p, size_t idx, int wb)
{
    switch (wb) {
    case 1: return ((const int8_t *)p)[idx];
    case 2: return ((const int16_t *)p)[idx];
    case 4: return ((const int32_t *)p)[idx];
    default: return ((const int64_t *)p)[idx];
    }
}
static void store_word(void *p, size_t idx, int wb, i128 v)
{
    switch (wb) {
    case 1: ((int8_t *)p)[idx] = (int8_t)v; break;
    case 2: ((int16_t *)p)[idx] = (int16_t)v; break;
    case 4: ((int32_t *)p)[idx] = (int32_t)v; break;
    default: ((int64_t *)p)[idx] = (int64_t)v; break;
    }
}
static int cword_bytes(qb_qdesc re, qb_qdesc im)
{
    int a = word_bytes_of(re), b = word_bytes_of(im);
    return a > b ? a : b;
}

/* ------------------------------- exported entry points ------------------------------- */

/* Scalar probes (used by the golden-vector tests). */
int64_t qbo_conv_i64(int64_t raw, int from_frac, const qb_qdesc *to) { return (int64_t)qbo_conv(raw, from_frac, *to); }
/* 128-bit input variant for the reference's wide known-answer tests (test/RND, test/TRN: 70.70 -> 1.1). */
int64_t qbo_conv_i128(int64_t hi, uint64_t lo, int from_frac, const qb_qdesc *to)
{
    i128 raw = ((i128)hi << 64) | (i128)lo;
    return (int64_t)qbo_conv(raw, from_frac, *to);
}
int64_t qbo_qmul_i64(int64_t a, const qb_qdesc *ta, int64_t b, const qb_qdesc *tb, const qb_qdesc *to)
{
    qval x = { a, *ta }, y = { b, *tb }; return (int64_t)qbo_qmul(x, y, *to).raw;
}
int64_t qbo_qadd_i64(int64_t a, const qb_qdesc *ta, int64_t b, const qb_qdesc *tb, const qb_qdesc *to)
{
    qval x = { a, *ta }, y = { b, *tb }; return (int64_t)qbo_qadd(x, y, *to).raw;
}
int64_t qbo_qsub_i64(int64_t a, const qb_qdesc *ta, int64_t b, const qb_qdesc *tb, const qb_qdesc *to)
{
    qval x = { a, *ta }, y = { b, *tb }; return (int64_t)qbo_qsub(x, y, *to).raw;
}

/* Type of a reduction result (depends on len and form, see tree_vec / tree_var). */
int qbo_reduce_out_type(const qb_reduce_plan *p, qb_qdesc *out)
{
    if (p->len == 0) return QB_EINVAL;
    *out = tree_type((size_t)p->len, p->elem, &p->levels, (int)p->tree_form, 0);
    out->word_bytes = 0;
    return QB_OK;
}

/* Qreduce over batch vectors.  partials: level-`partial_level` nodes of the VECTOR form
 * (ceil(len/2^partial_level) per vector, stored as int64), or NULL. */
int qbo_reduce_batched(const qb_reduce_plan *p, const void *x, void *out, int64_t *partials,
                       uint32_t partial_level, uint64_t batch)
{
    size_t n = (size_t)p->len;
    if (n == 0) return QB_EINVAL;
    int wb = word_bytes_of(p->elem);
    int64_t *xs = (int64_t *)malloc(n * sizeof(int64_t));
    if (!xs) return QB_ENOMEM;
    size_t per = n; for (unsigned l = 0; l < partial_level; l++) per = (per + 1) / 2;
    for (uint64_t b = 0; b < batch; b++) {
        for (size_t i = 0; i < n; i++) xs[i] = (int64_t)load_word(x, b * n + i, wb);
        qval r;
        if (p->tree_form == QB_TREE_VECTOR)
            r = tree_vec(xs, n, p->elem, &p->levels, partials ? partials + b * per : NULL, partial_level);
        else
            r = tree_var(xs, n, p->elem, &p->levels, 0);
        store_word(out, b, word_bytes_of(r.t), r.raw);
    }
    free(xs);
    return QB_OK;
}

/* Qgemul (readme.md:80-87): C[i,j] = conv_C( Qreduce<AddArgs>( Qmul<MulArgs>(opA(i,k), opB(k,j)) ...k ) ),
 * the final assignment through the converting ctor with C's modes (QuBLAS.h:2398-2411). */
int qbo_gemul_batched(const qb_gemul_plan *p, const void *A, const void *B, void *C, uint64_t batch)
{
    size_t M = p->M, N = p->N, K = p->K;
    if (!M || !N || !K) return QB_EINVAL;
    int wa = word_bytes_of(p->a), wbb = word_bytes_of(p->b), wc = word_bytes_of(p->c);
    int64_t *xs = (int64_t *)malloc(K * sizeof(int64_t));
    if (!xs) return QB_ENOMEM;
    for (uint64_t b = 0; b < batch; b++) {
        size_t oa = (size_t)(b * (uint64_t)p->stride_a), ob = (size_t)(b * (uint64_t)p->stride_b),
               oc = (size_t)(b * (uint64_t)p->stride_c);
        for (size_t j = 0; j < N; j++)
            for (size_t i = 0; i < M; i++) {
                for (size_t k = 0; k < K; k++) {
                    /* column-major, QuBLAS.h:2680-2692 */
                    size_t ia = p->trans_a ? (k + i * K) : (i + k * M);
                    size_t ib = p->trans_b ? (j + k * N) : (k + j * K);
                    qval a = { load_word(A, oa + ia, wa), p->a };
                    qval bb = { load_word(B, ob + ib, wbb), p->b };
                    xs[k] = (int64_t)qbo_qmul(a, bb, p->mul).raw;
                }
                qval r = tree(xs, K, p->mul, &p->add, p->tree_form);
                store_word(C, oc + i + j * M, wc, qbo_conv(r.raw, r.t.frac_bits, p->c));
            }
    }
    free(xs);
    return QB_OK;
}

/* TFComplexMul::mul, QuBLAS.h:3522-3534, on resolved node types. */
static cval tfcmul(cval x, cval y, const qb_tfc_types *t)
{
    qval a = x.re, b = x.im, c = y.re, d = y.im;
    qval Aq = qbo_qmul(qbo_qadd(a, b, t->ab), c, t->abc);                    /* :3524 */
    qval Bq = qbo_qmul(qbo_qadd(c, d, t->cd), b, t->bad);                    /* :3525 */
    qval Cq = qbo_qmul(qbo_qsub(b, a, t->ba), d, t->cdb);                    /* :3526 */
    cval r;
    r.re = qbo_qsub(Aq, Bq, t->AB);                                          /* :3528 */
    r.im = qbo_qsub(Bq, Cq, t->BC);                                          /* :3529 */
    return r;
}

/* BasicComplexMul::mul, QuBLAS.h:3437-3445, on resolved node types. */
static cval bcmul(cval x, cval y, const qb_bcm_types *t)
{
    qval a = x.re, b = x.im, c = y.re, d = y.im;
    cval r;
    r.re = qbo_qsub(qbo_qmul(a, c, t->ac), qbo_qmul(b, d, t->bd), t->acbd);      /* :3439 */
    r.im = qbo_qadd(qbo_qmul(a, d, t->ad), qbo_qmul(b, c, t->bc), t->adbc);      /* :3440 */
    return r;
}

/* real (x) complex Qmul, QuBLAS.h:3600-3637: the real operand multiplies each component, each product with
 * its own target type.  real_first: x.re is the real operand (x.im unused); otherwise y.re is (y.im unused). */
static cval rcmul(cval x, cval y, const qb_bcm_types *t, int real_first)
{
    cval r;
    r.re = qbo_qmul(x.re, y.re, t->ac);                                                   /* :3612 / :3633 */
    r.im = real_first ? qbo_qmul(x.re, y.im, t->ad) : qbo_qmul(x.im, y.re, t->bc);        /* :3613 / :3634 */
    return r;
}

static cval cproduct(int kind, cval X, cval Y, const qb_tfc_types *tf, const qb_bcm_types *bcm)
{
    if (kind == QB_CPROD_BASIC) return bcmul(X, Y, bcm);
    if (kind == QB_CPROD_REAL_A) return rcmul(X, Y, bcm, 1);
    if (kind == QB_CPROD_REAL_B) return rcmul(X, Y, bcm, 0);
    return tfcmul(X, Y, tf);
}

/* operand loads: a real operand stores one word per element */
static cval load_cval(const void *buf, size_t e, int wb, int is_real, qb_qdesc re_t, qb_qdesc im_t)
{
    cval v;
    if (is_real) { v.re.raw = load_word(buf, e, wb); v.re.t = re_t; v.im.raw = 0; v.im.t = re_t; }
    else { v.re.raw = load_word(buf, 2 * e, wb); v.re.t = re_t; v.im.raw = load_word(buf, 2 * e + 1, wb); v.im.t = im_t; }
    return v;
}

static void cproduct_types(int kind, const qb_tfc_types *tf, const qb_bcm_types *bcm, qb_qdesc *re, qb_qdesc *im)
{
    if (kind == QB_CPROD_BASIC) { *re = bcm->acbd; *im = bcm->adbc; }
    else if (kind == QB_CPROD_REAL_A) { *re = bcm->ac; *im = bcm->ad; }
    else if (kind == QB_CPROD_REAL_B) { *re = bcm->ac; *im = bcm->bc; }
    else { *re = tf->AB; *im = tf->BC; }
}

int qbo_tfcmul_batched(const qb_tfcmul_plan *p, const void *x, const void *y, void *out, uint64_t batch)
{
    int rx = p->product == QB_CPROD_REAL_A, ry = p->product == QB_CPROD_REAL_B;
    int wx = rx ? word_bytes_of(p->x_re) : cword_bytes(p->x_re, p->x_im), wy = ry ? word_bytes_of(p->y_re) : cword_bytes(p->y_re, p->y_im);
    qb_qdesc tre, tim;
    cproduct_types((int)p->product, &p->tf, &p->bcm, &tre, &tim);
    int wo = cword_bytes(tre, tim);
    size_t n = (size_t)(p->len * batch);
    for (size_t i = 0; i < n; i++) {
        cval X = load_cval(x, i, wx, rx, p->x_re, p->x_im);
        cval Y = load_cval(y, i, wy, ry, p->y_re, p->y_im);
        cval r = cproduct((int)p->product, X, Y, &p->tf, &p->bcm);
        store_word(out, 2 * i, wo, r.re.raw);
        store_word(out, 2 * i + 1, wo, r.im.raw);
    }
    return QB_OK;
}

/* Complex Qgemul: TFComplexMul products accumulated with complex Qadd = component-wise Qadd with the
 * level type (QuBLAS.h:3549-3568), then component-wise conversion into C's element type. */
int qbo_cgemul_batched(const qb_cgemul_plan *p, const void *A, const void *B, void *C, uint64_t batch)
{
    size_t M = p->M, N = p->N, K = p->K;
    if (!M || !N || !K) return QB_EINVAL;
    int ra = p->product == QB_CPROD_REAL_A, rb = p->product == QB_CPROD_REAL_B;
    int wa = ra ? word_bytes_of(p->a_re) : cword_bytes(p->a_re, p->a_im), wbb = rb ? word_bytes_of(p->b_re) : cword_bytes(p->b_re, p->b_im);
    int wc = cword_bytes(p->c_re, p->c_im);
    qb_qdesc tre, tim;
    cproduct_types((int)p->product, &p->tf, &p->bcm, &tre, &tim);
    int64_t *xr = (int64_t *)malloc(K * sizeof(int64_t)), *xi = (int64_t *)malloc(K * sizeof(int64_t));
    if (!xr || !xi) { free(xr); free(xi); return QB_ENOMEM; }
    for (uint64_t b = 0; b < batch; b++) {
        size_t oa = (size_t)(b * (uint64_t)p->stride_a), ob = (size_t)(b * (uint64_t)p->stride_b),
               oc = (size_t)(b * (uint64_t)p->stride_c);
        for (size_t j = 0; j < N; j++)
            for (size_t i = 0; i < M; i++) {
                for (size_t k = 0; k < K; k++) {
                    size_t ia = oa + (p->trans_a ? (k + i * K) : (i + k * M));
                    size_t ib = ob + (p->trans_b ? (j + k * N) : (k + j * K));
                    cval X = load_cval(A, ia, wa, ra, p->a_re, p->a_im);
                    cval Y = load_cval(B, ib, wbb, rb, p->b_re, p->b_im);
                    cval r = cproduct((int)p->product, X, Y, &p->tf, &p->bcm);
                    xr[k] = (int64_t)r.re.raw; xi[k] = (int64_t)r.im.raw;
                }
                qval rr = tree(xr, K, tre, &p->add_re, p->tree_form);
                qval ri = tree(xi, K, tim, &p->add_im, p->tree_form);
                size_t ic = oc + i + j * M;
                store_word(C, 2 * ic, wc, qbo_conv(rr.raw, rr.t.frac_bits, p->c_re));
                store_word(C, 2 * ic + 1, wc, qbo_conv(ri.raw, ri.t.frac_bits, p->c_im));
            }
    }
    free(xr); free(xi);
    return QB_OK;
}

int qbo_convert_batched(const qb_convert_plan *p, const void *in, void *out, uint64_t batch)
{
    int wi = word_bytes_of(p->from), wo = word_bytes_of(p->to);
    size_t n = (size_t)(p->len * batch);
    for (size_t i = 0; i < n; i++)
        store_word(out, i, wo, qbo_conv(load_word(in, i, wi), p->from.frac_bits, p->to));
    return QB_OK;
}

/* Counter-based synthetic raw words; must match the device generator bit for bit. */
static uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
int64_t qbo_synth_word(uint64_t seed, uint64_t stream_id, uint64_t index, int64_t lo, int64_t hi)
{
    uint64_t h = mix64(mix64(seed * 0xD1342543DE82EF95ull + stream_id) ^ index);
    uint64_t span = (uint64_t)(hi - lo) + 1;       /* span == 0 means the full 64-bit range */
    if (span == 0) return (int64_t)h;
    /* multiply-high range reduction: floor(h * span / 2^64) */
    uint64_t r = (uint64_t)(((unsigned __int128)h * span) >> 64);
    return lo + (int64_t)r;
}
int qbo_synth_fill(void *dst, uint64_t count, uint32_t word_bytes, int64_t lo, int64_t hi,
                   uint64_t seed, uint64_t stream_id, uint64_t first_index)
{
    for (uint64_t i = 0; i < count; i++)
        store_word(dst, (size_t)i, (int)word_bytes, qbo_synth_word(seed, stream_id, first_index + i, lo, hi));
    return QB_OK;
}

/* ---------------------------------------------------------------------------------------------
 * the steps on either side of the path in an accuracy sweep.
 * double -> Qu: Qu_s(double), QuBLAS.h:2387-2393 = the exact value (loadFromDouble into a 2400-bit
 * buffer, :664) rounded by the type's QuMode (fracConvert<1200+frac -> frac>) and overflowed by its
 * OfMode.  Restated on the 53-bit mantissa: raw = ovf(round(m * 2^e * 2^frac)).
 * ------------------------------------------------------------------------------------------- */
#include <math.h>
int64_t qbo_from_double(double v, const qb_qdesc *t)
{
    if (v == 0.0 || v != v || isinf(v)) return 0;                            /* :667-671 */
    int e = 0;
    double m = frexp(v, &e);                                                  /* v = m * 2^e, 0.5 <= |m| < 1 */
    i128 mi = (i128)(long long)ldexp(m, 53);                                  /* exact 53-bit signed mantissa */
    int sh = 53 - e - t->frac_bits;                                           /* raw = mi * 2^-sh */
    i128 r;
    if (sh >= 0) r = sh > 120 ? ((t->qu_mode == QB_TRN_TCPL && mi < 0) ? -1 : 0) : qbo_round(mi, sh, t->qu_mode);
    else if (-sh > 70) r = mi > 0 ? ((i128)1 << 124) : -((i128)1 << 124);   /* beyond every <= 64-bit type; low 64 bits are zero */
    else r = shl128(mi, -sh);                                                 /* |mi| < 2^53: exact below 2^123 */
    return (int64_t)qbo_ovf(r, *t);
}
int qbo_quantize_f64(const qb_qdesc *t, const double *in, void *out, uint64_t n)
{
    int wb = word_bytes_of(*t);
    for (uint64_t i = 0; i < n; i++) store_word(out, (size_t)i, wb, qbo_from_double(in[i], t));
    return QB_OK;
}
/* toDouble, QuBLAS.h:2413-2416 */
int qbo_dequantize_f64(const qb_qdesc *t, const void *in, double *out, uint64_t n)
{
    int wb = word_bytes_of(*t);
    for (uint64_t i = 0; i < n; i++) out[i] = ldexp((double)(int64_t)load_word(in, (size_t)i, wb), -t->frac_bits);
    return QB_OK;
}
/* error of a fixed-point result against a float64 reference: sums for SQNR and the max abs error */
int qbo_error_stats(const qb_qdesc *t, const void *words, const double *ref, uint64_t n, double *sum_sq_err, double *sum_sq_ref,
                    double *max_abs_err)
{
    int wb = word_bytes_of(*t);
    double se = 0, sr = 0, mx = 0;
    for (uint64_t i = 0; i < n; i++) {
        double v = ldexp((double)(int64_t)load_word(words, (size_t)i, wb), -t->frac_bits);
        double d = v - ref[i];
        se += d * d; sr += ref[i] * ref[i];
        if (fabs(d) > mx) mx = fabs(d);
    }
    *sum_sq_err = se; *sum_sq_ref = sr; *max_abs_err = mx;
    return QB_OK;
}

/* element-wise Qmul / Qadd / Qsub of two tensors (the reference's lazy expressions, QuBLAS.h:3783-3811, evaluated per element) */
int qbo_elementwise_batched(const qb_ew_plan *p, const void *a, const void *b, void *out, uint64_t batch)
{
    int wa = word_bytes_of(p->a), wb = word_bytes_of(p->b), wo = word_bytes_of(p->to);
    size_t n = (size_t)(p->len * batch);
    for (size_t i = 0; i < n; i++) {
        qval x = { load_word(a, i, wa), p->a }, y = { load_word(b, i, wb), p->b };
        qval r = p->op == QB_EW_MUL ? qbo_qmul(x, y, p->to) : (p->op == QB_EW_ADD ? qbo_qadd(x, y, p->to) : qbo_qsub(x, y, p->to));
        store_word(out, i, wo, r.raw);
    }
    return QB_OK;
}

/* ANUS::Qpoly (QuBLAS.h:4836-4853): Qpoly<a0>(x) = a0; Qpoly<a0, a1, ...>(x) = Qadd<A0>(a0, Qmul<A0>(x, Qpoly<a1, ...>(x))),
 * A0 = decltype(a0).  Evaluated innermost first. */
static qval poly_eval(const qb_poly_segment *s, qval x)
{
    qval acc = { s->coef[s->ncoef - 1].raw, s->coef[s->ncoef - 1].type };
    for (int i = (int)s->ncoef - 2; i >= 0; i--) {
        qval a = { s->coef[i].raw, s->coef[i].type };
        acc = qbo_qadd(a, qbo_qmul(x, acc, a.t), a.t);
    }
    return acc;
}

/* ANUS::Qapprox (QuBLAS.h:4866-4888): the first Segment with x.toDouble() < breakpoint, the last one otherwise; the value is
 * converted into decltype(x).  approx == 0: plain Qpoly of segment 0. */
int qbo_poly_batched(const qb_poly_plan *p, const void *x, void *out, uint64_t batch)
{
    if (p->nseg < 1 || p->nseg > QB_POLY_MAX_SEG || (!p->approx && p->nseg != 1)) return QB_EINVAL;
    qb_qdesc ot = p->approx ? p->x : p->seg[0].coef[0].type;
    int wx = word_bytes_of(p->x), wo = word_bytes_of(ot);
    size_t n = (size_t)(p->len * batch);
    for (size_t i = 0; i < n; i++) {
        qval xv = { load_word(x, i, wx), p->x };
        uint32_t s = 0;
        if (p->approx) {
            double xd = ldexp((double)(int64_t)xv.raw, -p->x.frac_bits);                 /* toDouble, :2413-2416 */
            while (s + 1 < p->nseg && !(xd < p->seg[s].breakpoint)) s++;
        }
        qval r = poly_eval(&p->seg[s], xv);
        if (p->approx) r.raw = qbo_conv(r.raw, r.t.frac_bits, p->x);                     /* decltype(x){ ... }, converting ctor :2398-2411 */
        store_word(out, i, wo, r.raw);
    }
    return QB_OK;
}
