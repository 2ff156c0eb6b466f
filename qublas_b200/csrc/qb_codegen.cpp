#include "qb_codegen.h"
#include <cstdio>
#include <cstring>
#include <sstream>

namespace qb {

std::string emit_qstage(const QStage &q)
{
    char buf[256];
    auto lit = [](long long v, char *o) {
        if (v == INT64_MIN) snprintf(o, 40, "(-9223372036854775807LL-1)"); else snprintf(o, 40, "%lldLL", v);
    };
    char lo[40], hi[40]; lit(q.lo, lo); lit(q.hi, hi);
    snprintf(buf, sizeof(buf), "qb::QStage{%d,%u,%u,%u,%u,%u,%u,%u,%u,{0,0,0,0,0,0,0},%s,%s}", (int)q.shift, q.qu, q.of, q.need_lo, q.need_hi,
             q.wrap_bits, q.wrap_signed, q.identity, q.wide, lo, hi);
    return buf;
}

std::string emit_addstage(const AddStage &a)
{
    char buf[64];
    snprintf(buf, sizeof(buf), "qb::AddStage{%u,%u,%u,{0,0,0,0,0},", a.shl_a, a.shl_b, a.sub);
    return std::string(buf) + emit_qstage(a.q) + "}";
}

static std::string emit_levels(const char *fn, const TreeProg &t, unsigned n)
{
    std::ostringstream o;
    o << "  static constexpr __host__ __device__ qb::AddStage " << fn << "(int l) {\n    constexpr qb::AddStage t[" << (n ? n : 1) << "] = {\n";
    for (unsigned l = 0; l < (n ? n : 1); l++) o << "      " << emit_addstage(t.lvl[l]) << ",\n";
    o << "    };\n    return t[l];\n  }\n";
    return o.str();
}

static std::string emit_flush(const char *fn, const TreeProg &t);

// ANUS polynomial program as a constexpr initialiser; only the segments and steps in use are spelled out (the rest is
// value-initialised), so the text stays short
std::string emit_polyprog(const PolyProg &P)
{
    std::ostringstream o;
    o << "qb::PolyProg{" << P.nseg << "u," << P.approx << "u,{\n";
    for (uint32_t s = 0; s < P.nseg; s++) {
        const PolySeg &g = P.seg[s];
        auto lit = [](int64_t v) { return v == INT64_MIN ? std::string("(-9223372036854775807LL-1)") : std::to_string((long long)v) + "LL"; };
        o << "      qb::PolySeg{" << lit(g.below) << "," << g.ncoef << "u,0u," << lit(g.top) << ",{";
        for (uint32_t i = 0; i + 1 < g.ncoef; i++)
            o << "\n        qb::PolyStep{" << emit_qstage(g.step[i].mul) << "," << emit_addstage(g.step[i].add) << "," << lit(g.step[i].a) << "},";
        o << "}," << emit_qstage(g.back) << "},\n";
    }
    o << "    }}";
    return o.str();
}
static std::string emit_epilogue(const PolyProg *epi, const PolyProg *pro_a = nullptr, const PolyProg *pro_b = nullptr)
{
    PolyProg z; memset(&z, 0, sizeof(z));
    std::ostringstream o;
    o << "  static constexpr bool EPI = " << (epi ? "true" : "false") << ", PRO_A = " << (pro_a ? "true" : "false") << ", PRO_B = " << (pro_b ? "true" : "false") << ";\n";
    o << "  static constexpr __host__ __device__ qb::PolyProg poly() { return " << emit_polyprog(epi ? *epi : z) << "; }\n";
    o << "  static constexpr __host__ __device__ qb::PolyProg poly_a() { return " << emit_polyprog(pro_a ? *pro_a : z) << "; }\n";
    o << "  static constexpr __host__ __device__ qb::PolyProg poly_b() { return " << emit_polyprog(pro_b ? *pro_b : z) << "; }\n";
    return o.str();
}
// batch strides of the plan (in elements), baked into the tag: a kernel is specialised per plan anyway
static std::string emit_strides(int64_t sa, int64_t sb, int64_t sc)
{
    std::ostringstream o;
    o << "  static constexpr unsigned long long SA = " << (unsigned long long)sa << "ULL, SB = " << (unsigned long long)sb << "ULL, SC = " << (unsigned long long)sc << "ULL;\n";
    return o.str();
}
static unsigned ilog2(uint64_t n) { unsigned l = 0; while (((uint64_t)1 << (l + 1)) <= n) l++; return l; }

static const char *word_t(int wb) { return wb == 1 ? "int8_t" : wb == 2 ? "int16_t" : wb == 4 ? "int32_t" : "int64_t"; }

static bool fp_mul_ok(const QStage &q, const TVal &x, const TVal &y);

// Does the register-tile template run Qmul<MulArgs> of this plan as one FFMA (qb_device.cuh:fp_mul_apply)?
bool gemul_fp_mul(const qb_gemul_plan &p, const GemulProg &g)
{
    const int s = g.mul.shift;
    const Ival ra = type_range(p.a), rb = type_range(p.b);
    const i128 c4[4] = { ra.lo * rb.lo, ra.lo * rb.hi, ra.hi * rb.lo, ra.hi * rb.hi };
    i128 lo = c4[0], hi = c4[0];
    for (int i = 1; i < 4; i++) { if (c4[i] < lo) lo = c4[i]; if (c4[i] > hi) hi = c4[i]; }
    const bool mode_ok = g.mul.qu == RND_CONV || g.mul.qu == TRN_TCPL;
    const i128 mag = (hi > -lo ? hi : -lo);
    const bool ops_ok = width(p.a) <= 24 && width(p.b) <= 24;
    const bool of_ok = g.mul.of == SAT_TCPL || g.mul.of == SAT_SMGN;
    return s >= 1 && s <= 30 && mode_ok && ops_ok && of_ok && !g.mul.wide && mag <= ((i128)1 << (22 + s));
}

// Conservative product thresholds for the per-instance clamp elision (qb_fast_gemul.cuh): an instance whose extreme
// corner products stay inside [p_lo, p_hi] cannot reach Qmul's saturation bounds, because
// round(p) <= floor(p / 2^s) + 1 and round(p) >= floor(p / 2^s).
void gemul_clamp_thresholds(const GemulProg &g, long long *p_hi, long long *p_lo)
{
    const int s = g.mul.shift;
    if (s > 0) { *p_hi = (long long)(((i128)g.mul.hi << s) - 1); *p_lo = (long long)((i128)g.mul.lo << s); }
    else       { *p_hi = g.mul.hi >> (-s); *p_lo = -((-(g.mul.lo)) >> (-s)); }
}

// Shared memory of the register-tile template (qb_fast_gemul.cuh:GemulTile) for `group` instances per CTA: a two-stage
// ring of staged instances, plus -- when the kernel makes a pass over every staged instance (FP32 conversion, operand
// ranges for the clamp elision, packed words) and the budget allows -- a separate 32-bit copy the tiles read (*sep).
// Packed int16 words cannot do without the copy: 0 is returned when it does not fit.
uint64_t gemul_tile_smem(const qb_gemul_plan &p, const GemulProg &g, int group, bool *sep)
{
    const uint64_t words = (uint64_t)(group > 0 ? group : 1) * ((uint64_t)p.M * p.K + (uint64_t)p.K * p.N);
    const bool narrow = g.a_wb == 2;
    const bool dyn = (g.mul.need_lo || g.mul.need_hi) && (g.mul.of == SAT_TCPL || g.mul.of == SAT_SMGN);
    const bool pass = narrow || dyn || gemul_fp_mul(p, g);
    const uint64_t stage = words * (uint64_t)g.a_wb, limit = 220 * 1024;
    const bool s = pass && stage + words * 4 + 64 <= limit;           // one staging buffer + the copy
    if (narrow && !s) return 0;
    if (sep) *sep = s;
    const uint64_t total = s ? stage + words * 4 + 64 : 2 * stage + 64;   // ... or the two-stage ring, converted in place
    return total <= limit ? total : 0;
}

std::string emit_gemul_cfg(const std::string &name, const qb_gemul_plan &p, const GemulProg &g, int group, const PolyProg *epi, const PolyProg *pro_a, const PolyProg *pro_b)
{
    std::ostringstream o;
    o << "struct " << name << " {\n";
    o << "  typedef int32_t I;\n";
    o << "  typedef " << word_t(g.a_wb) << " WA; typedef " << word_t(g.b_wb) << " WB; typedef " << word_t(g.c_wb) << " WC;\n";
    unsigned nlev_c = 0;
    while (((uint64_t)1 << nlev_c) < p.K) nlev_c++;                  // ceil(log2 K); == log2 K for the power-of-two case
    o << "  static constexpr int M = " << p.M << ", N = " << p.N << ", K = " << p.K << ", NLEV = " << nlev_c << ", NFLUSH = " << g.tree.nflush << ", G = " << (group > 0 ? group : 1) << ";\n";
    o << "  static constexpr bool TRANS_A = " << (p.trans_a ? "true" : "false") << ", TRANS_B = " << (p.trans_b ? "true" : "false") << ";\n";
    o << emit_strides(p.stride_a, p.stride_b, p.stride_c);
    o << "  static constexpr __host__ __device__ qb::QStage mul() { return " << emit_qstage(g.mul) << "; }\n";
    o << "  static constexpr __host__ __device__ qb::QStage out() { return " << emit_qstage(g.out) << "; }\n";
    o << emit_levels("lvl", g.tree, nlev_c);
    o << emit_flush("flush", g.tree);
    // FP32-FMA product quantizer: fma(a, b, 1.5 * 2^(23+s)) is exact-product-plus-one-IEEE-rounding at a
    // granularity of 2^s, i.e. Qmul's rounded right shift by s in ONE instruction on the FMA pipe
    // (round-to-nearest-even == RND::CONV, round toward -inf == TRN::TCPL; the half-up / half-down modes
    // would need C +- 2^(s-1), which is below the ulp of C and not representable, so they stay on the
    // integer path).  Legal iff the operands are exact floats and |a*b| stays within the binade
    // [2^(23+s), 2^(24+s)] around the magic constant.
    {
        const int s = g.mul.shift;
        const bool fp = gemul_fp_mul(p, g);
        long long p_hi, p_lo;
        gemul_clamp_thresholds(g, &p_hi, &p_lo);
        o << "  static constexpr long long P_HI = " << p_hi << "LL, P_LO = " << p_lo << "LL;\n";
        o << "  static constexpr bool FP_MUL = " << (fp ? "true" : "false") << ";\n";
        o << "  static constexpr int FP_SHIFT = " << (s > 0 ? s : 1) << ";\n";
        bool sep = false;
        gemul_tile_smem(p, g, group, &sep);
        o << "  static constexpr bool SEP = " << (sep ? "true" : "false") << ";\n";
        const Ival ra = type_range(p.a), rb = type_range(p.b);           // operand types' own bounds (32-bit register class: they fit)
        o << "  static constexpr long long A_LO = " << (long long)ra.lo << "LL, A_HI = " << (long long)ra.hi << "LL, B_LO = " << (long long)rb.lo << "LL, B_HI = " << (long long)rb.hi << "LL;\n";
    }
    o << emit_epilogue(epi, pro_a, pro_b);
    o << "};\n";
    return o.str();
}

// Tag struct for the thread-per-output template (qb_fast_gemul_thread.cuh): any shape, both forms, transposes.
std::string emit_gemul_thread_cfg(const std::string &name, const qb_gemul_plan &p, const GemulProg &g, int bits, const PolyProg *epi, const PolyProg *pro_a, const PolyProg *pro_b)
{
    std::ostringstream o;
    unsigned nlev = 0;
    while (((uint64_t)1 << nlev) < p.K) nlev++;                       // ceil(log2 K): levels 0 .. nlev-1 exist
    o << "struct " << name << " {\n";
    o << "  typedef " << (bits == 32 ? "int32_t" : "int64_t") << " I;\n";
    o << "  typedef " << word_t(g.a_wb) << " WA; typedef " << word_t(g.b_wb) << " WB; typedef " << word_t(g.c_wb) << " WC;\n";
    o << "  static constexpr unsigned M = " << p.M << ", N = " << p.N << ", K = " << p.K << ";\n";
    o << "  static constexpr int NLEV = " << nlev << ", NFLUSH = " << g.tree.nflush << ", TOP_SLOT = " << g.tree.top_slot << ";\n";
    o << "  static constexpr bool TRANS_A = " << (p.trans_a ? "true" : "false") << ", TRANS_B = " << (p.trans_b ? "true" : "false") << ";\n";
    o << emit_strides(p.stride_a, p.stride_b, p.stride_c);
    o << "  static constexpr __host__ __device__ qb::QStage mul() { return " << emit_qstage(g.mul) << "; }\n";
    o << "  static constexpr __host__ __device__ qb::QStage out() { return " << emit_qstage(g.out) << "; }\n";
    o << emit_levels("lvl", g.tree, nlev);
    o << "  static constexpr __host__ __device__ qb::FlushStep flush(int i) {\n    constexpr qb::FlushStep t[" << (g.tree.nflush ? g.tree.nflush : 1) << "] = {\n";
    for (unsigned i = 0; i < (g.tree.nflush ? g.tree.nflush : 1); i++) {
        const FlushStep &f = g.tree.flush[i];
        o << "      qb::FlushStep{" << (unsigned)f.kind << "," << (unsigned)f.slot << ",{0,0,0,0,0,0}," << emit_addstage(f.a) << "},\n";
    }
    o << "    };\n    return t[i];\n  }\n";
    o << emit_epilogue(epi, pro_a, pro_b);
    o << "};\n";
    return o.str();
}

static std::string emit_flush(const char *fn, const TreeProg &t)
{
    std::ostringstream o;
    o << "  static constexpr __host__ __device__ qb::FlushStep " << fn << "(int i) {\n    constexpr qb::FlushStep t[" << (t.nflush ? t.nflush : 1) << "] = {\n";
    for (unsigned i = 0; i < (t.nflush ? t.nflush : 1); i++) {
        const FlushStep &f = t.flush[i];
        o << "      qb::FlushStep{" << (unsigned)f.kind << "," << (unsigned)f.slot << ",{0,0,0,0,0,0}," << emit_addstage(f.a) << "},\n";
    }
    o << "    };\n    return t[i];\n  }\n";
    return o.str();
}

static std::string emit_tfprog(const TfProg &t)
{
    return "qb::TfProg{" + emit_addstage(t.ab) + ",\n      " + emit_addstage(t.cd) + ",\n      " + emit_addstage(t.ba) + ",\n      " + emit_qstage(t.A) + ",\n      " +
           emit_qstage(t.B) + ",\n      " + emit_qstage(t.C) + ",\n      " + emit_addstage(t.re) + ",\n      " + emit_addstage(t.im) + "}";
}

static std::string emit_bcmprog(const BcmProg &t)
{
    return "qb::BcmProg{" + emit_qstage(t.ac) + ",\n      " + emit_qstage(t.bd) + ",\n      " + emit_qstage(t.ad) + ",\n      " + emit_qstage(t.bc) + ",\n      " +
           emit_addstage(t.re) + ",\n      " + emit_addstage(t.im) + "}";
}

// Tag struct for the complex thread-per-output template (qb_fast_gemul_thread.cuh): any product kind, shape, form.
std::string emit_cgemul_thread_cfg(const std::string &name, const qb_cgemul_plan &p, const CGemulProg &g, int bits)
{
    std::ostringstream o;
    unsigned nlev = 0;
    while (((uint64_t)1 << nlev) < p.K) nlev++;
    o << "struct " << name << " {\n";
    o << "  typedef " << (bits == 32 ? "int32_t" : "int64_t") << " I;\n";
    o << "  typedef " << word_t(g.a_wb) << " WA; typedef " << word_t(g.b_wb) << " WB; typedef " << word_t(g.c_wb) << " WC;\n";
    o << "  static constexpr unsigned M = " << p.M << ", N = " << p.N << ", K = " << p.K << ";\n";
    o << "  static constexpr int KIND = " << g.kind << ", NLEV = " << nlev << ", NFLUSH_RE = " << g.tree_re.nflush << ", NFLUSH_IM = " << g.tree_im.nflush
      << ", TOP_RE = " << g.tree_re.top_slot << ", TOP_IM = " << g.tree_im.top_slot << ";\n";
    o << "  static constexpr bool TRANS_A = " << (p.trans_a ? "true" : "false") << ", TRANS_B = " << (p.trans_b ? "true" : "false") << ";\n";
    o << emit_strides(p.stride_a, p.stride_b, p.stride_c);
    o << "  static constexpr __host__ __device__ qb::TfProg tf() { return " << emit_tfprog(g.tf) << "; }\n";
    o << "  static constexpr __host__ __device__ qb::BcmProg bcm() { return " << emit_bcmprog(g.bcm) << "; }\n";
    o << "  static constexpr __host__ __device__ qb::QStage out_re() { return " << emit_qstage(g.out_re) << "; }\n";
    o << "  static constexpr __host__ __device__ qb::QStage out_im() { return " << emit_qstage(g.out_im) << "; }\n";
    o << emit_levels("lvl_re", g.tree_re, nlev) << emit_levels("lvl_im", g.tree_im, nlev);
    o << emit_flush("flush_re", g.tree_re) << emit_flush("flush_im", g.tree_im);
    {
        // TFComplexMul products that are exact as one FFMA (used by the register-tile kernel, qb_fast_cgemul_tile.cuh)
        CgemulFp f;
        if ((g.kind == QB_CPROD_TF || g.kind == QB_CPROD_BASIC) && bits == 32) f = cgemul_fp(p, g);
        o << "  static constexpr bool FP_A = " << (f.fa ? "true" : "false") << ", FP_B = " << (f.fb ? "true" : "false") << ", FP_C = " << (f.fc ? "true" : "false") << ";\n";
        o << "  static constexpr int SH_A = " << (f.fa ? (int)g.tf.A.shift : 1) << ", SH_B = " << (f.fb ? (int)g.tf.B.shift : 1) << ", SH_C = " << (f.fc ? (int)g.tf.C.shift : 1) << ";\n";
        o << "  static constexpr bool FP_BCM = " << (f.fbcm ? "true" : "false") << ";\n";
        o << "  static constexpr int SH_AC = " << (f.fbcm ? (int)g.bcm.ac.shift : 1) << ", SH_BD = " << (f.fbcm ? (int)g.bcm.bd.shift : 1) << ", SH_AD = " << (f.fbcm ? (int)g.bcm.ad.shift : 1)
          << ", SH_BC = " << (f.fbcm ? (int)g.bcm.bc.shift : 1) << ";\n";
    }
    o << "};\n";
    return o.str();
}

// Is fma(x, y, 1.5 * 2^(23+s)) an exact implementation of stage q on operands with these ranges?
static bool fp_mul_ok(const QStage &q, const TVal &x, const TVal &y)
{
    const int s = q.shift;
    if (s < 1 || s > 30) return false;
    if (!(q.qu == RND_CONV || q.qu == TRN_TCPL) || !(q.of == SAT_TCPL || q.of == SAT_SMGN)) return false;
    auto fits24 = [](const Ival &r) { return r.lo >= -((i128)1 << 24) && r.hi <= ((i128)1 << 24); };
    if (!fits24(x.r) || !fits24(y.r)) return false;                                   // exact floats
    const i128 c4[4] = { x.r.lo * y.r.lo, x.r.lo * y.r.hi, x.r.hi * y.r.lo, x.r.hi * y.r.hi };
    i128 mag = 0;
    for (i128 c : c4) { i128 a = c < 0 ? -c : c; if (a > mag) mag = a; }
    if (mag > ((i128)1 << (22 + s))) return false;                                    // stays in the binade of the constant
    const i128 lo = (i128)q.lo, hi = (i128)q.hi;                                       // biased clamp bounds must be ints
    return lo >= -((i128)1 << 30) && hi <= ((i128)1 << 30);
}

// Which nodes of the TFComplexMul dataflow the mat-vec rows template runs on the FP32 pipe (qb_fast_cgemv.cuh).
CgemulFp cgemul_fp(const qb_cgemul_plan &p, const CGemulProg &g)
{
    CgemulFp f;
    if (g.kind == QB_CPROD_BASIC) {
        // BasicComplexMul (QuBLAS.h:3437-3445): four products of raw operands
        const TVal a = { p.a_re, type_range(p.a_re) }, b = { p.a_im, type_range(p.a_im) }, c = { p.b_re, type_range(p.b_re) }, d = { p.b_im, type_range(p.b_im) };
        f.fbcm = fp_mul_ok(g.bcm.ac, a, c) && fp_mul_ok(g.bcm.bd, b, d) && fp_mul_ok(g.bcm.ad, a, d) && fp_mul_ok(g.bcm.bc, b, c);
        return f;
    }
    // FP32-FMA quantizer for the three TF products where it is exact
    Analysis an; TVal re, im, mid[3];
    TVal a = { p.a_re, type_range(p.a_re) }, b = { p.a_im, type_range(p.a_im) }, c = { p.b_re, type_range(p.b_re) }, d = { p.b_im, type_range(p.b_im) };
    build_tf(a, b, c, d, p.tf, &re, &im, an, mid);
    f.fa = fp_mul_ok(g.tf.A, mid[0], c); f.fb = fp_mul_ok(g.tf.B, mid[1], b); f.fc = fp_mul_ok(g.tf.C, mid[2], d);
    // The pre-adds a + b and b - a feed those products as floats.  When the stage is an exact add (no alignment,
    // no rounding shift) whose only other effect is a SAT::TCPL clamp, and both operands are exact floats below 2^23,
    // the add itself runs on the FP pipe (FADD is exact below 2^24, the clamp is min/max on floats) and one
    // int->float conversion per product disappears.
    auto fadd_ok = [](const AddStage &s, const TVal &x, const TVal &y) {
        auto small = [](const Ival &r) { return r.lo >= -((i128)1 << 23) && r.hi <= ((i128)1 << 23); };
        return !s.shl_a && !s.shl_b && s.q.shift == 0 && !s.q.wrap_bits && !s.q.wide && (s.q.of == SAT_TCPL || (!s.q.need_lo && !s.q.need_hi)) &&
               small(x.r) && small(y.r) && s.q.lo >= -(1LL << 24) && s.q.hi <= (1LL << 24);
    };
    f.fadd_ab = f.fa && fadd_ok(g.tf.ab, a, b);
    f.fadd_ba = f.fc && fadd_ok(g.tf.ba, b, a);
    f.ab = mid[0].r; f.cd = mid[1].r; f.ba = mid[2].r;
    return f;
}

std::string emit_cgemul_cfg(const std::string &name, const qb_cgemul_plan &p, const CGemulProg &g)
{
    std::ostringstream o;
    o << "struct " << name << " {\n";
    o << "  typedef int32_t I;\n";
    o << "  typedef " << word_t(g.a_wb) << " WA; typedef " << word_t(g.b_wb) << " WB; typedef " << word_t(g.c_wb) << " WC;\n";
    o << "  static constexpr int M = " << p.M << ", N = " << p.N << ", K = " << p.K << ", NLEV = " << ilog2(p.K) << ";\n";
    o << emit_strides(p.stride_a, p.stride_b, p.stride_c);
    o << "  static constexpr __host__ __device__ qb::TfProg tf() { return qb::TfProg{" << emit_addstage(g.tf.ab) << ",\n      "
      << emit_addstage(g.tf.cd) << ",\n      " << emit_addstage(g.tf.ba) << ",\n      " << emit_qstage(g.tf.A) << ",\n      "
      << emit_qstage(g.tf.B) << ",\n      " << emit_qstage(g.tf.C) << ",\n      " << emit_addstage(g.tf.re) << ",\n      "
      << emit_addstage(g.tf.im) << "}; }\n";
    o << "  static constexpr __host__ __device__ qb::QStage out_re() { return " << emit_qstage(g.out_re) << "; }\n";
    o << "  static constexpr __host__ __device__ qb::QStage out_im() { return " << emit_qstage(g.out_im) << "; }\n";
    o << emit_levels("lvl_re", g.tree_re, ilog2(p.K));
    o << emit_levels("lvl_im", g.tree_im, ilog2(p.K));
    {
        const CgemulFp f = cgemul_fp(p, g);
        o << "  static constexpr bool FP_A = " << (f.fa ? "true" : "false") << ", FP_B = " << (f.fb ? "true" : "false") << ", FP_C = " << (f.fc ? "true" : "false") << ";\n";
        o << "  static constexpr int SH_A = " << (f.fa ? (int)g.tf.A.shift : 1) << ", SH_B = " << (f.fb ? (int)g.tf.B.shift : 1) << ", SH_C = " << (f.fc ? (int)g.tf.C.shift : 1) << ";\n";
        o << "  static constexpr bool FADD_AB = " << (f.fadd_ab ? "true" : "false") << ", FADD_BA = " << (f.fadd_ba ? "true" : "false") << ";\n";
    }
    o << "};\n";
    return o.str();
}

std::string emit_reduce_cfg(const std::string &name, const qb_reduce_plan &p, const ReduceProg &r)
{
    std::ostringstream o;
    unsigned logn = ilog2(p.len);
    o << "struct " << name << " {\n  typedef int32_t I;\n  static constexpr int LOGN = " << logn << ";\n";
    o << emit_levels("lvl", r.tree, logn);
    o << "};\n";
    return o.str();
}

// Tag struct for the any-length vector-form reduction (qb_fast_reduce.cuh, reduce_any_body): level and carry tables
// over all ceil(log2 len) levels.
std::string emit_reduce_any_cfg(const std::string &name, const qb_reduce_plan &p, const ReduceProg &r)
{
    std::ostringstream o;
    unsigned nlev = 0;
    while (((uint64_t)1 << nlev) < p.len) nlev++;
    o << "struct " << name << " {\n  typedef int32_t I;\n  static constexpr int NLEVELS = " << nlev << ";\n";
    o << emit_levels("lvl", r.tree, nlev);
    o << "  static constexpr __host__ __device__ qb::QStage carry(int l) {\n    constexpr qb::QStage t[" << (nlev ? nlev : 1) << "] = {\n";
    for (unsigned l = 0; l < (nlev ? nlev : 1); l++) o << "      " << emit_qstage(r.tree.carry[l]) << ",\n";
    o << "    };\n    return t[l];\n  }\n";
    o << "};\n";
    return o.str();
}

// Tag struct for the element-wise templates (qb_fast_ew.cuh).  kind >= 0: complex product of that kind (uses tf / bcm);
// kind < 0: real element-wise op `op` (0 mul, 1 add, 2 sub, 3 convert) with `stage`.
std::string emit_ew_cfg(const std::string &name, int kind, const TfProg *tf, const BcmProg *bcm, int op, const AddStage *stage, int xw, int yw, int ow, int bits)
{
    std::ostringstream o;
    TfProg ztf; BcmProg zb; AddStage zs;
    memset(&ztf, 0, sizeof(ztf)); memset(&zb, 0, sizeof(zb)); memset(&zs, 0, sizeof(zs));
    o << "struct " << name << " {\n";
    o << "  typedef " << (bits == 32 ? "int32_t" : "int64_t") << " I;\n";
    o << "  typedef " << word_t(xw) << " WX; typedef " << word_t(yw) << " WY; typedef " << word_t(ow) << " WO;\n";
    o << "  static constexpr int KIND = " << kind << ", OP = " << op << ";\n";
    o << "  static constexpr __host__ __device__ qb::TfProg tf() { return " << emit_tfprog(tf ? *tf : ztf) << "; }\n";
    o << "  static constexpr __host__ __device__ qb::BcmProg bcm() { return " << emit_bcmprog(bcm ? *bcm : zb) << "; }\n";
    o << "  static constexpr __host__ __device__ qb::AddStage stage() { return " << emit_addstage(stage ? *stage : zs) << "; }\n";
    o << "};\n";
    return o.str();
}

// Tag struct of the specialised ANUS kernel (qb_fast_ew.cuh:poly_body)
std::string emit_poly_cfg(const std::string &name, const PolyProg &prog, int xw, int ow, int bits)
{
    std::ostringstream o;
    o << "struct " << name << " {\n";
    o << "  typedef " << (bits == 32 ? "int32_t" : "int64_t") << " I;\n";
    o << "  typedef " << word_t(xw) << " WX; typedef " << word_t(xw) << " WY; typedef " << word_t(ow) << " WO;\n";
    o << emit_epilogue(&prog);
    o << "};\n";
    return o.str();
}

std::string emit_plan_bytes(const void *plan, size_t n)
{
    std::ostringstream o;
    const unsigned char *b = (const unsigned char *)plan;
    o << "{";
    for (size_t i = 0; i < n; i++) { o << (unsigned)b[i]; if (i + 1 < n) o << ","; if ((i & 63) == 63) o << "\n"; }
    o << "}";
    return o.str();
}

}  // namespace qb
