// ref_bench_c4.cpp — the reference's OWN CPU implementation of BASELINE config 4: batched 256x256
// complex mat-vec with TFComplexMul products.  TEST / BASELINE INFRASTRUCTURE, compiled against the unmodified header
// where it lies (oracle/Makefile, target `ref`); nothing from the reference is copied into this repo.
//
// The variadic complex Qreduce over 256 products does not finish compiling, so each output row is the
// header-true composition of SURVEY.md 8(c): Qmul<TFComplexMul<...>> per product (QuBLAS.h:3510-3535), then two REAL
// vector-form reductions, over the real and over the imaginary parts (QuBLAS.h:4986-5001) -- complex Qadd is
// component-wise with the same level type (:3555-3558), and for a power-of-two K both tree forms coincide.
#include "QuBLAS.h"
#include <cstdint>
using namespace QuBLAS;

namespace c4 {
using t = Qu<intBits<3>, fracBits<12>>;
using P = Qu<intBits<6>, fracBits<12>>;
using S = Qu<intBits<7>, fracBits<12>>;
using cx = Qcomplex<t, t>;
using tf = TFComplexMul<abT<Qu<intBits<4>, fracBits<12>>>, cdT<Qu<intBits<4>, fracBits<11>>>, abcT<P>, cdbT<P>, badT<P>, ABT<S>, BCT<S>>;
template <int l> using lv = Qu<intBits<8 + l>, fracBits<12 - l>, QuMode<RND::CONV>>;
using list = TypeList<lv<0>, lv<1>, lv<2>, lv<3>, lv<4>, lv<5>, lv<6>, lv<7>>;
constexpr int M = 256, K = 256;
static inline cx mk(const int32_t *p) { cx q; q.real.data.data = p[0]; q.imag.data.data = p[1]; return q; }
}  // namespace c4

// H: n_inst x (256 x 256) complex, column-major, AoS {re, im} int32; x: n_inst x 256 complex; y: n_inst x 256 complex
extern "C" void qref_c4_matvec(const int32_t *H, const int32_t *x, int32_t *y, int n_inst)
{
    using namespace c4;
    using prod_t = decltype(Qmul<tf>(cx(), cx()));
    static Qu<dim<K>, typename prod_t::realType> vr;
    static Qu<dim<K>, typename prod_t::imagType> vi;
    for (int b = 0; b < n_inst; b++) {
        const int32_t *Hb = H + (size_t)b * M * K * 2, *xb = x + (size_t)b * K * 2;
        int32_t *yb = y + (size_t)b * M * 2;
        for (int i = 0; i < M; i++) {
            for (int j = 0; j < K; j++) {
                auto p = Qmul<tf>(mk(Hb + ((size_t)j * M + i) * 2), mk(xb + (size_t)j * 2));
                vr[j] = p.real; vi[j] = p.imag;
            }
            lv<7> yr = Qreduce<list>(vr), yi = Qreduce<list>(vi);
            yb[2 * i] = yr.data.data; yb[2 * i + 1] = yi.data.data;
        }
    }
}
