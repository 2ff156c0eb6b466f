// ref_bench.cpp — the reference's OWN CPU implementation of BASELINE config 3 (ref_bench_c2.cpp / ref_bench_c4.cpp: configs 2, 4).
// TEST / BASELINE INFRASTRUCTURE.  Compiled (oracle/Makefile, target `ref`) against the unmodified
// header where it lies (/root/reference/include/QuBLAS.h); the output oracle/_ref/libqref_bench.so
// is git-ignored and travels to the GPU box.  Nothing from the reference is copied into this repo.
//
// Qgemul exists only in the reference's readme (readme.md:80-87), so the timed code is the
// composition a user of the header would write: Qmul<MulArgs> per product, variadic
// Qreduce<TypeList<...>> over the K products (QuBLAS.h:4953-4957), converting assignment into C's
// element type (QuBLAS.h:2398-2411).
#include "QuBLAS.h"
#include <cstdint>
#include <utility>
using namespace QuBLAS;

namespace c3 {   // BASELINE config 3: batched 64x64x64, 16-bit signed operands
using in_t  = Qu<intBits<7>, fracBits<8>>;
using mul_t = Qu<intBits<14>, fracBits<8>, QuMode<RND::CONV>, OfMode<SAT::TCPL>>;
template <int l> using lv = Qu<intBits<15 + l>, fracBits<8 - l>, QuMode<RND::POS_INF>, OfMode<SAT::TCPL>>;
using list  = TypeList<lv<0>, lv<1>, lv<2>, lv<3>, lv<4>, lv<5>>;
using c_t   = Qu<intBits<17>, fracBits<4>, QuMode<RND::POS_INF>, OfMode<SAT::TCPL>>;
constexpr int M = 64, N = 64, K = 64;

static inline in_t mk(int32_t raw) { in_t q; q.data.data = raw; return q; }

template <size_t... Ks>
static inline int32_t dot(const int32_t *A, const int32_t *B, int i, int j, std::index_sequence<Ks...>)
{
    c_t cij = Qreduce<list>(Qmul<mul_t>(mk(A[i + Ks * M]), mk(B[Ks + j * K]))...);
    return cij.data.data;
}
}  // namespace c3

extern "C" void qref_c3_gemul(const int32_t *A, const int32_t *B, int32_t *C, int n_inst)
{
    using namespace c3;
    for (int b = 0; b < n_inst; b++) {
        const int32_t *a = A + (size_t)b * M * K, *bb = B + (size_t)b * K * N;
        int32_t *c = C + (size_t)b * M * N;
        for (int j = 0; j < N; j++)
            for (int i = 0; i < M; i++) c[i + j * M] = dot(a, bb, i, j, std::make_index_sequence<K>{});
    }
}

