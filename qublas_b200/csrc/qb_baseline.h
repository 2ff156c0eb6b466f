// qb_baseline.h — the plans that get an ahead-of-time specialised kernel (BASELINE.json configs,
// exact types from SURVEY.md 8(d)).  Shared by the build-time generator (qb_gen) and the registry.
#pragma once
#include <cstring>
#include "../../include/qb200.h"

namespace qb {

inline qb_qdesc qd(int i, int f, bool s = true, int q = QB_TRN_TCPL, int o = QB_SAT_TCPL)
{
    qb_qdesc d; d.int_bits = (int16_t)i; d.frac_bits = (int16_t)f; d.is_signed = s; d.qu_mode = (uint8_t)q; d.of_mode = (uint8_t)o; d.word_bytes = 0;
    return d;
}

// Config 3: batched 64x64x64, A,B : Qu<7,8>; MulArgs Qu<14,8,RND::CONV,SAT::TCPL>;
// AddArgs TypeList<Qu<15+l, 8-l, RND::POS_INF, SAT::TCPL>, l=0..5>; C : Qu<17,4,RND::POS_INF,SAT::TCPL>
inline qb_gemul_plan baseline_c3()
{
    qb_gemul_plan p; memset(&p, 0, sizeof(p));
    p.a = p.b = qd(7, 8);
    p.mul = qd(14, 8, true, QB_RND_CONV, QB_SAT_TCPL);
    p.add.n = 6;
    for (int l = 0; l < 6; l++) p.add.lv[l] = qd(15 + l, 8 - l, true, QB_RND_POS_INF, QB_SAT_TCPL);
    p.c = qd(17, 4, true, QB_RND_POS_INF, QB_SAT_TCPL);
    p.M = p.N = p.K = 64;
    p.stride_a = p.stride_b = p.stride_c = 64 * 64;
    return p;
}

// Config 4: batched 256x256 complex mat-vec, t = Qu<3,12>; TFComplexMul<abT<4.12>, cdT<4.11>,
// abcT/cdbT/badT<6.12>, ABT/BCT<7.12>> (b-a: default merger 3.12); tree TypeList<Qu<8+l,12-l,RND::CONV>, l=0..7>
inline qb_cgemul_plan baseline_c4()
{
    qb_cgemul_plan p; memset(&p, 0, sizeof(p));
    qb_qdesc t = qd(3, 12), P = qd(6, 12), S = qd(7, 12);
    p.a_re = p.a_im = p.b_re = p.b_im = t;
    p.tf.ab = qd(4, 12); p.tf.cd = qd(4, 11); p.tf.ba = t;
    p.tf.abc = p.tf.bad = p.tf.cdb = P; p.tf.AB = p.tf.BC = S;
    p.add_re.n = p.add_im.n = 8;
    for (int l = 0; l < 8; l++) p.add_re.lv[l] = p.add_im.lv[l] = qd(8 + l, 12 - l, true, QB_RND_CONV, QB_SAT_TCPL);
    p.c_re = p.c_im = p.add_re.lv[7];
    p.M = 256; p.N = 1; p.K = 256;
    p.tree_form = QB_TREE_VARIADIC;
    p.stride_a = 256 * 256; p.stride_b = 256; p.stride_c = 256;
    return p;
}

// Config 2: Qreduce of 2^24 x Qu<8,8>; list "a" = 8 entries Qu<9+l,8-l,M,SAT::TCPL> (last reused for levels 8..23),
// list "b" = 24 entries, Qu<17+ceil((l-7)/2),0,M,SAT::TCPL> for l = 8..23; M = any of the 7 QuModes.
inline qb_reduce_plan baseline_c2(bool list_b, int mode)
{
    qb_reduce_plan p; memset(&p, 0, sizeof(p));
    p.elem = qd(8, 8);
    p.levels.n = list_b ? 24 : 8;
    for (int l = 0; l < 8; l++) p.levels.lv[l] = qd(9 + l, 8 - l, true, mode, QB_SAT_TCPL);
    if (list_b) for (int l = 8; l < 24; l++) p.levels.lv[l] = qd(17 + (l - 7 + 1) / 2, 0, true, mode, QB_SAT_TCPL);
    p.len = 1ull << 24;
    p.tree_form = QB_TREE_VECTOR;
    return p;
}

}  // namespace qb
