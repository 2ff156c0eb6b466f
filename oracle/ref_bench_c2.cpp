// ref_bench_c2.cpp — the reference's OWN CPU implementation of BASELINE config 2: Qreduce of a
// 2^24-element Qu<8,8> vector with a per-level TypeList.  TEST / BASELINE INFRASTRUCTURE, compiled against the unmodified
// header where it lies (oracle/Makefile, target `ref`); nothing from the reference is copied into this repo.
//
// The header's vector-form Qreduce (QuBLAS.h:4960-4990) only instantiates for <= 1000 elements (std::get on the
// std::array storage, :2853-2863), so 2^24 = 512 * 512 * 64 is reduced by three CHAINED header reductions: blocks of 512
// with TypeList<L0..L8>, their 512 partials with the list shifted by nine levels, the 64 partials of that with the list
// shifted by eighteen.  For a power-of-two length this is the same tree (separability).
#include "QuBLAS.h"
#include <cstdint>
using namespace QuBLAS;

namespace c2 {
using e_t = Qu<intBits<8>, fracBits<8>>;
// C2b: 24 level types, L_l = Qu<9+l, 8-l> for l < 8, Qu<17 + ceil((l-7)/2), 0> above (no saturation: the result depends on every level)
template <class M, int l> using lb = Qu<intBits<(l < 8 ? 9 + l : 17 + (l - 6) / 2)>, fracBits<(l < 8 ? 8 - l : 0)>, QuMode<M>, OfMode<SAT::TCPL>>;
// C2a: 8 level types, the last one (16.1) reused for levels 8..23 (saturates)
template <class M, int l> using la = Qu<intBits<9 + (l < 8 ? l : 7)>, fracBits<8 - (l < 8 ? l : 7)>, QuMode<M>, OfMode<SAT::TCPL>>;

template <template <class, int> class L, class M>
static int32_t chain(const int32_t *x, int32_t *level9 /* 2^15 partials or nullptr */)
{
    using s1 = TypeList<L<M, 0>, L<M, 1>, L<M, 2>, L<M, 3>, L<M, 4>, L<M, 5>, L<M, 6>, L<M, 7>, L<M, 8>>;
    using s2 = TypeList<L<M, 9>, L<M, 10>, L<M, 11>, L<M, 12>, L<M, 13>, L<M, 14>, L<M, 15>, L<M, 16>, L<M, 17>>;
    using s3 = TypeList<L<M, 18>, L<M, 19>, L<M, 20>, L<M, 21>, L<M, 22>, L<M, 23>>;
    static Qu<dim<512>, e_t> v0;
    static Qu<dim<512>, L<M, 8>> v1;
    static Qu<dim<64>, L<M, 17>> v2;
    for (int b2 = 0; b2 < 64; b2++) {
        for (int b1 = 0; b1 < 512; b1++) {
            const int32_t *p = x + ((size_t)b2 * 512 + b1) * 512;
            for (int i = 0; i < 512; i++) v0[i].data.data = p[i];
            L<M, 8> r = Qreduce<s1>(v0);
            v1[b1] = r;
            if (level9) level9[b2 * 512 + b1] = r.data.data;
        }
        L<M, 17> r1 = Qreduce<s2>(v1);
        v2[b2] = r1;
    }
    L<M, 23> r2 = Qreduce<s3>(v2);
    return r2.data.data;
}
}  // namespace c2

// variant: 0 = C2a (8-entry list), 1 = C2b (24-entry list); mode: the reference's QuMode ids (0 POS_INF, 4 CONV, 5 TCPL).
// Returns 0 and the root in *out; -1 for a combination that is not instantiated.
extern "C" int qref_c2_chain(const int32_t *x, int variant, int mode, int32_t *out, int32_t *level9)
{
    using namespace c2;
    if (variant == 1 && mode == 0) { *out = chain<lb, RND::POS_INF>(x, level9); return 0; }
    if (variant == 1 && mode == 4) { *out = chain<lb, RND::CONV>(x, level9); return 0; }
    if (variant == 1 && mode == 5) { *out = chain<lb, TRN::TCPL>(x, level9); return 0; }
    if (variant == 0 && mode == 0) { *out = chain<la, RND::POS_INF>(x, level9); return 0; }
    return -1;
}

// one 512-block stage (levels 0..8, C2a/C2b agree there), kept for the round-1 callers
extern "C" void qref_c2_reduce512(const int32_t *x, int32_t *out, int n_blocks)
{
    using namespace c2;
    using list = TypeList<lb<RND::POS_INF, 0>, lb<RND::POS_INF, 1>, lb<RND::POS_INF, 2>, lb<RND::POS_INF, 3>, lb<RND::POS_INF, 4>,
                          lb<RND::POS_INF, 5>, lb<RND::POS_INF, 6>, lb<RND::POS_INF, 7>>;
    static Qu<dim<512>, e_t> v;
    for (int b = 0; b < n_blocks; b++) {
        for (int i = 0; i < 512; i++) v[i].data.data = x[(size_t)b * 512 + i];
        auto r = Qreduce<list>(v);
        out[b] = r.data.data;
    }
}
