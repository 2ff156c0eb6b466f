// The reference's readme.md:20-127 usage example, compiled UNCHANGED against this repo's QuBLAS.h
// (minus the ANUS LUT lines, which are outside the hot path; the BitStream vector is given all six
// of its values — with the readme's four the reference itself throws std::invalid_argument).  Prints the raw words of
// every result so that the test can compare them with the reference's own answers.
// Exit code 3 = no CUDA device (scalar parts still ran).
#include "QuBLAS.h"
#include <cstdio>
#include <iostream>

using namespace QuBLAS;

int main()
{
    Qu<intBits<8>, fracBits<8>, isSigned<true>, QuMode<TRN::TCPL>, OfMode<SAT::ZERO>> a = 1.0;

    using type1 = Qu<isSigned<true>, intBits<6>, fracBits<3>, OfMode<SAT::ZERO>>;
    type1 q1 = 1.0;

    using type2 = Qu<intBits<6>, fracBits<-3>>;
    type2 q2 = 16;

    using vecType = Qu<dim<4>, intBits<4>>;
    vecType v1 = {1.0, 2.0, 3.0, 4.0};

    using matType = Qu<dim<4, 4>, type1>;
    matType m1 = {1.0, 2.0, 3.0, 4.0,
                  5.0, 6.0, 7.0, 8.0,
                  9.0, 10.0, 11.0, 12.0,
                  13.0, 14.0, 15.0, 16.0};

    using highDimType = Qu<dim<4, 1, 2>, type1>;
    highDimType h1 = {1.0, 2.0, 3.0, 4.0,
                      5.0, 6.0, 7.0, 8.0};

    auto elem = m1[1, 2];
    std::printf("a %d q1 %d q2 %d elem %d h1[3,0,1] %d\n", (int)a.data.data, (int)q1.data.data, (int)q2.data.data, (int)elem.data.data,
                (int)h1[3, 0, 1].data.data);

    // manually provided inputs: header-only variadic form
    auto red3 = Qreduce<type1>(q1, q2, q1, q2);
    std::printf("red3 %d\n", (int)red3.data.data);

    using list = TypeList<type1, type2>;
    using c_t_1 = Qcomplex<type1, type2>;
    using c_t_2 = Qu<type1, type2>;
    static_assert(std::is_same_v<c_t_1, c_t_2>);
    c_t_1 complex1 = {1, 2};
    c_t_2 complex2 = {3, 4};
    auto cprod = Qmul<TFComplexMul<abT<type2>, cdT<type1>, baT<type2>, badT<type2>, ABT<type2>, BCT<type2>>>(complex1, complex2);
    std::printf("cprod %d %d\n", (int)cprod.real.data.data, (int)cprod.imag.data.data);

    // BitStream (readme.md:111-127): host-only wire format
    using vec_t_bits = Qu<dim<6>, type1>;
    vec_t_bits vec_bits = {1.0, 2.0, 3.0, 4.0, 5.0, 6.0};
    auto res = BitStream<r2l<3>, r2l<2>>(vec_bits);
    std::cout << "bits " << res << std::endl;
    auto z = BitStream<vec_t_bits, r2l<3>, r2l<2>>(res);
    std::printf("bitsz");
    for (std::size_t i = 0; i < 6; i++) std::printf(" %d", (int)z[i].data.data);
    std::printf("\n");
    std::fflush(stdout);

    try {
        // Tree-based reductions on tensors -> GPU
        auto red1 = Qreduce<type2>(v1);
        auto red2 = Qreduce<list>(m1);
        std::printf("red1 %d red2 %d\n", (int)red1.data.data, (int)red2.data.data);

        matType m3;
        Qgemul<QgemulAddArgs<list>,    // Tree-based reduction for dot product
               QgemulMulArgs<type1>,   // Specify quantization mode for multiplication
               QgemulTransposedA<true> // Tags in any order and are all optional as well
               >(m3, m1, m1);
        std::printf("m3");
        for (std::size_t i = 0; i < 16; i++) std::printf(" %d", (int)m3[i].data.data);
        std::printf("\n");

        matType m4;
        Qgemul<QgemulMulArgs<type1>, QgemulAddArgs<list>>(m4, m1, m1);
        std::printf("m4");
        for (std::size_t i = 0; i < 16; i++) std::printf(" %d", (int)m4[i].data.data);
        std::printf("\n");

        using complex_vec_t = Qu<dim<3>, c_t_1>;
        complex_vec_t complex_vec = {a, a, a};
        auto complex_vec_mul = Qmul<TFComplexMul<abT<type2>, cdT<type1>, baT<type2>, badT<type2>, ABT<type2>, BCT<type2>>>(complex_vec, complex_vec);
        std::printf("cvec");
        for (std::size_t i = 0; i < 3; i++) std::printf(" %d %d", (int)complex_vec_mul[i].real.data.data, (int)complex_vec_mul[i].imag.data.data);
        std::printf("\n");
    } catch (const std::runtime_error &e) {
        std::printf("runtime_error: %s\n", e.what());
        return 3;
    }
    return 0;
}
