// ANUS::Qpoly / Qapprox (reference include/QuBLAS.h:4829-4897) on scalars.  Compiled against the
// reference header by `oracle/ref_harness_gen.py --text` (golden tests/golden/ref_anus.txt) and
// against this repo's header; outputs must be identical.  Coefficient types avoid RND::CONV: the
// reference's double constructor is defective for that mode (parity domain).
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

    for (int raw = -4096; raw < 4096; raw += 37) {
        x_t x;
        x.fill(raw);
        auto p1 = ANUS::Qpoly<a0>(x);
        auto p2 = ANUS::Qpoly<a0, a1>(x);
        auto p3 = ANUS::Qpoly<a0, a1, a2>(x);
        auto p4 = ANUS::Qpoly<b0, b1, b2, a3>(x);
        auto q = ANUS::Qapprox<ANUS::Segment<-2.0, a0, a1>, ANUS::Segment<0.5, b0, b1, b2>, ANUS::Segment<3.0, a3, a2>>(x);
        auto r = ANUS::Qapprox_s<ANUS::Segment<1.0, b1>, ANUS::Segment<100.0, a1, b2>>::execute(x);
        std::cout << raw << " " << (long long)p1.data.data << " " << (long long)p2.data.data << ":" << decltype(p2)::intB << "." << decltype(p2)::fracB
                  << " " << (long long)p3.data.data << " " << (long long)p4.data.data << ":" << decltype(p4)::intB << "." << decltype(p4)::fracB
                  << " " << (long long)q.data.data << " " << (long long)r.data.data << "\n";
    }
    return 0;
}
