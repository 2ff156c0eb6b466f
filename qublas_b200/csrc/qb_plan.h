// qb_plan.h — host-side plan compiler: qb200.h plans -> device programs (qb_prog.h).
//
// Mirrors the reference's compile-time type rules (QuBLAS.h:3079-3140 mergers, :4913 level
// selection) at run time and adds what a GPU kernel needs on top: exact interval analysis of every
// intermediate (which register width is safe, which clamps are provably dead, whether the whole
// accumulation is lossless) and the parity-domain fence of DESIGN.md / SURVEY.md 2.3.
#pragma once
#include <string>
#include "../../include/qb200.h"
#include "qb_prog.h"

namespace qb {

typedef __int128 i128;

struct Ival { i128 lo, hi; };
struct TVal { qb_qdesc t; Ival r; int frac() const { return t.frac_bits; } };

struct Analysis {
    int  max_bits = 1;        // widest exact intermediate incl. rounding bias (two's complement bits)
    int  max_reg_bits = 1;    // widest value that lives BETWEEN stages (leaves and stage outputs)
    int  ref_width = 1;       // widest STATIC ArbiInt width the reference would instantiate
    int  max_shift = 0;       // largest shift amount any stage uses (left, right or alignment)
    bool lossless = true;     // every quantizer so far is an identity
    int  status = 0;          // QB_OK / QB_EINVAL / QB_EDOMAIN
    std::string err;
    void fail(int code, const std::string &m) { if (!status) { status = code; err = m; } }
};

int  natural_word(qb_qdesc t);                 // 4 / 8 (reference storage word, QuBLAS.h:353)
int  word_bytes(qb_qdesc t);                   // t.word_bytes or natural
int  width(qb_qdesc t);
Ival type_range(qb_qdesc t);
bool check_type(qb_qdesc t, Analysis &an, const char *what);
qb_qdesc add_merge_default(qb_qdesc a, qb_qdesc b);

QStage   make_stage(int from_frac, Ival in, int in_width, qb_qdesc to, TVal *out, Analysis &an);
QStage   make_mul(TVal a, TVal b, qb_qdesc to, TVal *out, Analysis &an);
AddStage make_add(TVal a, TVal b, qb_qdesc to, bool sub, TVal *out, Analysis &an);

// Tree over `count` leaves of type/interval `leaf`.  Returns the result's type/interval.
TVal build_tree(uint64_t count, TVal leaf, const qb_levels &L, int form, TreeProg *prog, Analysis &an);

TfProg build_tf(TVal a, TVal b, TVal c, TVal d, const qb_tfc_types &t, TVal *re, TVal *im, Analysis &an, TVal *mid = nullptr /* [3]: ab, cd, ba */);

struct ReduceProg {
    TreeProg tree;
    TVal     out;
    int      elem_wb, out_wb;
};
struct GemulProg {
    QStage   mul;
    TreeProg tree;
    QStage   out;          // conversion of the tree result into C's element type
    int      a_wb, b_wb, c_wb;
    TVal     prod, acc;
};
BcmProg build_bcm(TVal a, TVal b, TVal c, TVal d, const qb_bcm_types &t, TVal *re, TVal *im, Analysis &an);
BcmProg build_rc(int kind, TVal a, TVal b, TVal c, TVal d, const qb_bcm_types &t, TVal *re, TVal *im, Analysis &an);

struct CGemulProg {
    int      kind;         // enum qb_cproduct
    TfProg   tf;
    BcmProg  bcm;
    TreeProg tree_re, tree_im;
    QStage   out_re, out_im;
    int      a_wb, b_wb, c_wb;
};

int compile_reduce(const qb_reduce_plan &p, ReduceProg *out, Analysis &an);
int compile_gemul(const qb_gemul_plan &p, GemulProg *out, Analysis &an);
int compile_cgemul(const qb_cgemul_plan &p, CGemulProg *out, Analysis &an);
int compile_tfcmul(const qb_tfcmul_plan &p, TfProg *out, BcmProg *bout, TVal *re, TVal *im, Analysis &an);
// ANUS::Qpoly / Qapprox over a tensor (QuBLAS.h:4829-4897); *out_type = type of the result
int compile_poly(const qb_poly_plan &p, PolyProg *out, qb_qdesc *out_type, Analysis &an);
void sanitize32(PolyProg &t);

// Register class of a compiled plan: 32 iff every leaf and every stage output fits int32; else 64.
// Stages whose exact input is wider than 32 bits (e.g. a 17x16-bit product) are marked `wide` and
// evaluated in 64 bits inside an otherwise 32-bit kernel.
inline int arith_bits(const Analysis &an) { return an.max_reg_bits <= 32 ? 32 : 64; }
void sanitize32(TreeProg &t);
void sanitize32(TfProg &t);
void sanitize32(BcmProg &t);
void sanitize32_stage(QStage &q);

// tcgen05 kind::i8 legality: operands fit int8, the whole accumulation is an exact integer sum.
bool gemul_tensor_core_ok(const qb_gemul_plan &p, const GemulProg &g, const Analysis &an);

}  // namespace qb
