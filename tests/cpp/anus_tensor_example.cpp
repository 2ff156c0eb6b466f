// ANUS::Qpoly / Qapprox / Qtable on TENSORS (GPU, qb200_poly_batched / qb200_table_batched) against the scalar forms of the
// same header, which tests/golden/ref_anus.txt pins to the reference (include/QuBLAS.h:4829-4897).  Same coefficients as
// tests/cpp/anus_example.cpp.  Without a GPU the first tensor call throws (exit code 3): there is no CPU fallback.
#include "QuBLAS.h"
#include <iostream>

using namespace QuBLAS;

int main()
{
    using x_t = Qu<intBits<4>, fracBits<8>>;
    using c_t = Qu<intBits<3>, fracBits<6>, QuMode<RND::POS_INF>>;
    using w_t = Qu<intBits<2>, fracBits<10>, QuMode<TRN::SMGN>, OfMode<SAT::ZERO>>;
    constexpr c_t a0 = 0.5, a1 = -1.25, a2 = 0.375, a3 = 2.0;
    constexpr w_t b0 = -0.3333, b1 = 0.75, b2 = -0.0625;
    constexpr std::size_t N = 8192;
    using vec_t = Qu<dim<N>, x_t>;

    vec_t v;
    for (std::size_t i = 0; i < N; i++) v[i].fill((int)i - 4096);          // every raw value of x_t
    try {
        auto p3 = ANUS::Qpoly<a0, a1, a2>(v);
        auto p4 = ANUS::Qpoly<b0, b1, b2, a3>(v);
        auto q = ANUS::Qapprox<ANUS::Segment<-2.0, a0, a1>, ANUS::Segment<0.5, b0, b1, b2>, ANUS::Segment<3.0, a3, a2>>(v);
        auto r = ANUS::Qapprox_s<ANUS::Segment<1.0, b1>, ANUS::Segment<100.0, a1, b2>>::execute(v);
        auto t1 = ANUS::Qtable<ANUS::rsqrtFunc>(v);
        auto t2 = ANUS::Qtable<ANUS::reciprocalFunc>(v);
        auto t3 = ANUS::Qtable<ANUS::sqrtFunc>(v);
        static_assert(std::is_same_v<std::remove_cvref_t<decltype(p3[0])>, c_t> && std::is_same_v<std::remove_cvref_t<decltype(p4[0])>, w_t>);
        static_assert(std::is_same_v<std::remove_cvref_t<decltype(q[0])>, x_t> && std::is_same_v<std::remove_cvref_t<decltype(t1[0])>, x_t>);
        std::size_t bad = 0;
        for (std::size_t i = 0; i < N; i++) {
            const x_t x = v[i];
            bad += p3[i].data.data != ANUS::Qpoly<a0, a1, a2>(x).data.data;
            bad += p4[i].data.data != ANUS::Qpoly<b0, b1, b2, a3>(x).data.data;
            bad += q[i].data.data != ANUS::Qapprox<ANUS::Segment<-2.0, a0, a1>, ANUS::Segment<0.5, b0, b1, b2>, ANUS::Segment<3.0, a3, a2>>(x).data.data;
            bad += r[i].data.data != ANUS::Qapprox_s<ANUS::Segment<1.0, b1>, ANUS::Segment<100.0, a1, b2>>::execute(x).data.data;
            bad += t1[i].data.data != ANUS::Qtable<ANUS::rsqrtFunc>(x).data.data;
            bad += t2[i].data.data != ANUS::Qtable<ANUS::reciprocalFunc>(x).data.data;
            bad += t3[i].data.data != ANUS::Qtable<ANUS::sqrtFunc>(x).data.data;
        }
        x_t four; four.fill(4 * 256);
        std::cout << "tensor ANUS: " << (bad ? "MISMATCH" : "OK") << " (" << bad << " of " << 7 * N << ")\n";
        std::cout << "rsqrt(4.0) raw " << (long long)ANUS::Qtable<ANUS::rsqrtFunc>(four).data.data << " sqrt(4.0) raw " << (long long)ANUS::Qtable<ANUS::sqrtFunc>(four).data.data << "\n";
        return bad ? 1 : 0;
    } catch (const std::runtime_error &e) {
        std::cout << "runtime_error: " << e.what() << "\n";
        return 3;
    }
}
