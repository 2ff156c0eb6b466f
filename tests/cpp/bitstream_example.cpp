// BitStream wire format (the reference's RTL test-vector I/O, include/QuBLAS.h:4545-4827; main.cpp:6-43;
// readme.md:111-127).  The same source is compiled against the reference header (golden output,
// tests/golden/ref_bitstream.txt, written by oracle/ref_harness_gen.py --text) and against this
// repo's header; the outputs must be identical.
#include "QuBLAS.h"
#include <iostream>

using namespace QuBLAS;

int main()
{
    using type1 = Qu<isSigned<true>, intBits<6>, fracBits<3>, OfMode<SAT::ZERO>>;
    using utype = Qu<isSigned<false>, intBits<3>, fracBits<2>>;
    using vec_t = Qu<dim<6>, type1>;
    vec_t v;
    const int raws[6] = { 8, 16, 24, 32, -8, 511 };
    for (int i = 0; i < 6; i++) v[i].fill(raws[i]);

    auto s1 = BitStream<r2l<3>, r2l<2>>(v);
    auto s2 = BitStream<l2r, l2r>(v);
    auto s3 = BitStream<r2l<2>, l2r>(v);
    auto s4 = BitStream<l2r, r2l<5>>(v);
    std::cout << s1 << "\n" << s2 << "\n" << s3 << "\n" << s4 << "\n";

    auto z1 = BitStream<vec_t, r2l<3>, r2l<2>>(s1);
    auto z2 = BitStream<vec_t, l2r, l2r>(s2);
    auto z3 = BitStream<vec_t, r2l<2>, l2r>(s3);
    for (int i = 0; i < 6; i++) std::cout << (long long)z1[i].data.data << " " << (long long)z2[i].data.data << " " << (long long)z3[i].data.data << "\n";

    // separators are ignored when parsing
    auto z4 = BitStream<vec_t, l2r, l2r>(std::string("0000001000_0000010000 0000011000,0000100000|1111111000 0111111111"));
    for (int i = 0; i < 6; i++) std::cout << (long long)z4[i].data.data << " ";
    std::cout << "\n";

    // scalars
    type1 q; q.fill(-3);
    utype u; u.fill(21);
    std::cout << BitStream<l2r>(q) << " " << BitStream<r2l<2>>(q) << " " << BitStream<l2r>(u) << "\n";
    auto qb = BitStream<type1, l2r>(std::string("0000101101"));
    auto ub = BitStream<utype, l2r>(std::string("10101"));
    std::cout << (long long)qb.data.data << " " << (long long)ub.data.data << "\n";

    // complex elements: "(re, im)" per element on export; separators ignored on import
    using ctype = Qcomplex<type1, utype>;
    using cvec_t = Qu<dim<4>, ctype>;
    cvec_t cv;
    for (int i = 0; i < 4; i++) cv[i].fill(raws[i] + 1, i + 3);
    auto cs1 = BitStream<l2r, l2r>(cv);
    auto cs2 = BitStream<r2l<2>, l2r>(cv);
    std::cout << cs1 << "\n" << cs2 << "\n";
    auto cz1 = BitStream<cvec_t, l2r, l2r>(cs1);
    auto cz2 = BitStream<cvec_t, r2l<2>, l2r>(cs2);
    for (int i = 0; i < 4; i++)
        std::cout << (long long)cz1[i].real.data.data << "," << (long long)cz1[i].imag.data.data << " " << (long long)cz2[i].real.data.data << "," << (long long)cz2[i].imag.data.data << "\n";
    ctype cq; cq.fill(-2, 9);
    std::cout << BitStream<l2r>(cq) << "\n";
    auto cqb = BitStream<ctype, l2r>(std::string("(1111111110, 01001)"));
    std::cout << (long long)cqb.real.data.data << "," << (long long)cqb.imag.data.data << "\n";

    // a 4x4 matrix streams in storage (column-major) order
    using mat_t = Qu<dim<4, 4>, utype>;
    mat_t mt;
    for (int i = 0; i < 16; i++) mt[i].fill(i + 1);
    auto ms = BitStream<r2l<4>, r2l<>>(mt);
    std::cout << ms << "\n";
    auto mz = BitStream<mat_t, r2l<4>, r2l<>>(ms);
    for (int i = 0; i < 16; i++) std::cout << (long long)mz[i].data.data << " ";
    std::cout << "\n";

    // operator<< on tensors: doubles with 4 decimals; a 2-D tensor prints as aligned rows
    std::cout << v << "\n" << cv << "\n" << mt << "\n" << z4 << "\n";
    using mat2_t = Qu<dim<2, 3>, type1>;
    mat2_t m2 = { -1.5, 2.25, 30.0, -40.125, 0.5, 6.0 };
    std::cout << m2 << "\n" << m2[1, 2] << " " << cq << "\n";
    return 0;
}
