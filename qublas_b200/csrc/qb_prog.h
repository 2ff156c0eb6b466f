// qb_prog.h — "compiled" quantizer programs shared by the host plan compiler and the device kernels.
//
// A plan (include/qb200.h) is lowered on the host into a small POD program: one QStage per
// quantizer (fracConvert + intConvert pair, QuBLAS.h:2002-2334 of the reference), one AddStage per
// tree level (Qadd_s::add, QuBLAS.h:3177-3204) and a flush program for non-power-of-two counts
// (QuBLAS.h:4942-4949 variadic / :4977-4980 vector).  The structs are literal types with public
// members only, so one and the same kernel body can take them either
//   * as a __grid_constant__ kernel parameter (runtime-descriptor kernels: any plan), or
//   * as a C++20 class-type template argument (compile-time-specialised kernels: every shift,
//     bound and mode folds to an immediate, dead clamps disappear).
#pragma once
#include <stdint.h>

#define QBP_MAX_LEVELS 40      // 2^40 leaves
#define QBP_MAX_FLUSH  96

namespace qb {

enum : int { RND_POS_INF = 0, RND_NEG_INF = 1, RND_ZERO = 2, RND_INF = 3, RND_CONV = 4, TRN_TCPL = 5, TRN_SMGN = 6 };
enum : int { SAT_TCPL = 0, SAT_ZERO = 1, SAT_SMGN = 2, WRP_TCPL = 3 };

// v' = ovf(round(v, shift)): the converting step applied to an exact intermediate.
struct QStage {
    int8_t  shift;       // > 0: rounded right shift; <= 0: exact left shift by -shift
    uint8_t qu;          // QuMode id
    uint8_t of;          // OfMode id
    uint8_t need_lo;     // interval analysis: value can fall below lo  (SAT) / outside the wrap range (WRP)
    uint8_t need_hi;     // interval analysis: value can exceed hi
    uint8_t wrap_bits;   // WRP::TCPL: signed -> keep W bits sign-extended; unsigned -> keep W-1 bits
    uint8_t wrap_signed;
    uint8_t identity;    // 1 if the whole stage is provably v' = v << -shift with shift <= 0 and no clamp
    uint8_t wide;        // the exact input of this stage needs more than 32 bits (its output may still fit 32)
    uint8_t pad[7];
    int64_t lo, hi;      // SAT bounds of the target type (SMGN: symmetric lower bound)
};

// r = stage( (a << shl_a) +/- (b << shl_b) )
struct AddStage {
    uint8_t shl_a, shl_b;
    uint8_t sub;
    uint8_t pad[5];
    QStage  q;
};

enum : int { FL_LOAD = 0, FL_CONV = 1, FL_ADD = 2 };
// Flush program over the level-indexed stack: cur = st[slot] | cur = conv(cur) | cur = add(st[slot], cur)
struct FlushStep {
    uint8_t kind, slot;
    uint8_t pad[6];
    AddStage a;          // FL_CONV uses a.q only
};

// Adjacent-pair tree over `count` leaves consumed in index order with a level-indexed stack:
// after leaf k, every set bit l of (k+1) has a pending node in st[l].
struct TreeProg {
    uint32_t  nlev;                     // merge levels that can occur: floor(log2(count))
    uint32_t  nflush;                   // 0 when count is a power of two
    uint32_t  top_slot;                 // slot holding the result when nflush == 0
    uint32_t  pad;
    AddStage  lvl[QBP_MAX_LEVELS];      // lvl[l]: both operands are level-l nodes (same type); built up to ceil(log2(count))
    QStage    carry[QBP_MAX_LEVELS];    // vector form: odd leftover at level l converted into level l's type (QuBLAS.h:4977-4980)
    FlushStep flush[QBP_MAX_FLUSH];
};

// TFComplexMul dataflow (QuBLAS.h:3522-3534): x = a+bi, y = c+di
struct TfProg {
    AddStage ab, cd, ba;     // a+b, c+d, b-a
    QStage   A, B, C;        // (a+b)*c, (c+d)*b, (b-a)*d product quantizers
    AddStage re, im;         // A-B, B-C
};

// BasicComplexMul dataflow (QuBLAS.h:3437-3445): re = ac - bd, im = ad + bc
struct BcmProg {
    QStage   ac, bd, ad, bc;
    AddStage re, im;
};

// ANUS::Qpoly / Qapprox (QuBLAS.h:4829-4897): per segment a Horner chain acc = add(a_i, mul(x, acc)), i = n-2 .. 0,
// starting from acc = a_(n-1); Qapprox converts the value back into x's type.
#define QBP_POLY_MAX_COEF 8
#define QBP_POLY_MAX_SEG  8
struct PolyStep {
    QStage   mul;        // Qmul<A_i>(x, acc)
    AddStage add;        // Qadd<A_i>(a_i, product)
    int64_t  a;          // raw a_i
};
struct PolySeg {
    int64_t  below;      // the segment applies when raw(x) < below  (= ceil(breakpoint * 2^fracBits(x)))
    uint32_t ncoef, pad;
    int64_t  top;        // raw a_(n-1)
    PolyStep step[QBP_POLY_MAX_COEF - 1];
    QStage   back;       // Qapprox: converting ctor into x's type
};
struct PolyProg {
    uint32_t nseg, approx;
    PolySeg  seg[QBP_POLY_MAX_SEG];
};

}  // namespace qb
