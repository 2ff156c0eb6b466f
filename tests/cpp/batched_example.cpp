// C++ drop-in boundary, batched and complex entry points: every GPU result is compared with the
// header's own scalar (host, constexpr) arithmetic, which tests/test_host_header.py pins against the
// reference's goldens.  Prints "OK" lines; any mismatch returns 1.  Exit code 3 = no CUDA device.
#include "QuBLAS.h"
#include <cstdio>

using namespace QuBLAS;

using in_t  = Qu<intBits<7>, fracBits<8>>;
using mul_t = Qu<intBits<14>, fracBits<8>, QuMode<RND::CONV>, OfMode<SAT::TCPL>>;
template <int l> using lv = Qu<intBits<15 + l>, fracBits<8 - l>, QuMode<RND::POS_INF>, OfMode<SAT::TCPL>>;
using levels = TypeList<lv<0>, lv<1>, lv<2>, lv<3>>;
using c_t = Qu<intBits<17>, fracBits<4>, QuMode<RND::POS_INF>, OfMode<SAT::TCPL>>;

int main()
{
    int bad = 0;
    try {
        // ---- batched real Qgemul: 16x16x16, 40 instances in one call ----
        using A_t = Qu<dim<16, 16>, in_t>; using C_t = Qu<dim<16, 16>, c_t>;
        const std::size_t n = 40;
        Qbatch<A_t> A(n), B(n); Qbatch<C_t> C(n);
        A.synth(1, 0); B.synth(1, 1);
        QgemulBatched<QgemulAddArgs<levels>, QgemulMulArgs<mul_t>>(C, A, B);
        for (std::size_t inst : { (std::size_t)0, (std::size_t)17, n - 1 }) {
            A_t a = A.get(inst), b = B.get(inst); C_t c = C.get(inst);
            for (int i = 0; i < 16; i++) for (int j = 0; j < 16; j++) {
                mul_t p[16];
                for (int k = 0; k < 16; k++) p[k] = Qmul<mul_t>(a[i, k], b[k, j]);
                c_t want = Qreduce<levels>(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9], p[10], p[11], p[12], p[13], p[14], p[15]);
                if (want.data.data != c[i, j].data.data) bad++;
            }
        }
        std::printf("batched Qgemul: %s\n", bad ? "MISMATCH" : "OK");
        static_assert(sizeof(Qbatch<A_t>::raw_t) == 2 && sizeof(Qbatch<C_t>::raw_t) == 4, "narrowest storage word: 16-bit operands, 22-bit results");

        // ---- the same sweep device-resident (inputs generated in HBM, only statistics come back), and in the reference's words ----
        {
            Qbatch<A_t> dA(n, Qmem::device), dB(n, Qmem::device); Qbatch<C_t> dC(n, Qmem::device), hC(n);
            dA.synth(1, 0); dB.synth(1, 1);
            QgemulBatched<QgemulAddArgs<levels>, QgemulMulArgs<mul_t>>(dC, dA, dB);
            hC.copy_from(dC);
            int bad_dev = 0;
            for (std::size_t w = 0; w < hC.words(); w++) if (hC.raw()[w] != C.raw()[w]) bad_dev++;
            Qbatch<A_t, Qwords::reference> rA(n, Qmem::pageable), rB(n, Qmem::pageable); Qbatch<C_t, Qwords::reference> rC(n, Qmem::pageable);
            static_assert(sizeof(decltype(rA)::raw_t) == 4, "reference words: int32 for widths up to 32");
            rA.synth(1, 0); rB.synth(1, 1);
            QgemulBatched<QgemulAddArgs<levels>, QgemulMulArgs<mul_t>>(rC, rA, rB);
            for (std::size_t w = 0; w < rC.words(); w++) if (rC.raw()[w] != C.raw()[w]) bad_dev++;
            Qerror e_dev = QgemulError<QgemulAddArgs<levels>, QgemulMulArgs<mul_t>>(dC, dA, dB);
            Qerror e_host = QgemulError<QgemulAddArgs<levels>, QgemulMulArgs<mul_t>>(C, A, B);
            if (e_dev.n != n * 256 || e_dev.sum_sq_err != e_host.sum_sq_err || e_dev.max_abs_err != e_host.max_abs_err) bad_dev++;
            // the result type has 4 fraction bits, the tree drops bits at every level: the error is a few lsbs of C
            if (!(e_dev.max_abs_err > 0 && e_dev.max_abs_err < 64.0 / 16 && e_dev.sqnr_db() > 20)) bad_dev++;
            bool threw = false;
            try { (void)dC.get(0); } catch (const std::runtime_error &) { threw = true; }
            if (!threw) bad_dev++;
            std::printf("device-resident sweep: %s (SQNR %.1f dB, max |err| %.4f)\n", bad_dev ? "MISMATCH" : "OK", e_dev.sqnr_db(), e_dev.max_abs_err);
            bad += bad_dev;
        }

        // ---- complex Qgemul (8x8 times 8x1) with TFComplexMul and with the default BasicComplexMul ----
        using t = Qu<intBits<3>, fracBits<12>>;
        using cx = Qcomplex<t, t>;
        using P = Qu<intBits<6>, fracBits<12>>; using S = Qu<intBits<7>, fracBits<12>>;
        using TF = TFComplexMul<abT<Qu<intBits<4>, fracBits<12>>>, cdT<Qu<intBits<4>, fracBits<11>>>, abcT<P>, cdbT<P>, badT<P>, ABT<S>, BCT<S>>;
        using clv = TypeList<Qu<intBits<8>, fracBits<12>, QuMode<RND::CONV>>, Qu<intBits<9>, fracBits<11>, QuMode<RND::CONV>>, Qu<intBits<10>, fracBits<10>, QuMode<RND::CONV>>>;
        using out_s = Qu<intBits<10>, fracBits<10>, QuMode<RND::CONV>>;
        using H_t = Qu<dim<8, 8>, cx>; using x_t = Qu<dim<8>, cx>; using y_t = Qu<dim<8>, Qcomplex<out_s, out_s>>;
        H_t H; x_t x; y_t y, yb;
        for (std::size_t i = 0; i < 64; i++) { H[i].real.fill(); H[i].imag.fill(); }
        for (std::size_t i = 0; i < 8; i++) { x[i].real.fill(); x[i].imag.fill(); }
        Qgemul<QgemulAddArgs<clv>, QgemulMulArgs<TF>>(y, H, x);
        Qgemul<QgemulAddArgs<clv>>(yb, H, x);                                   // no MulArgs: BasicComplexMul<>
        int bad_tf = 0, bad_b = 0;
        for (int i = 0; i < 8; i++) {
            auto w = Qreduce<clv>(Qmul<TF>(H[i, 0], x[0]), Qmul<TF>(H[i, 1], x[1]), Qmul<TF>(H[i, 2], x[2]), Qmul<TF>(H[i, 3], x[3]),
                                  Qmul<TF>(H[i, 4], x[4]), Qmul<TF>(H[i, 5], x[5]), Qmul<TF>(H[i, 6], x[6]), Qmul<TF>(H[i, 7], x[7]));
            out_s wr = w.real, wi = w.imag;
            if (wr.data.data != y[i].real.data.data || wi.data.data != y[i].imag.data.data) bad_tf++;
            auto v = Qreduce<clv>(Qmul(H[i, 0], x[0]), Qmul(H[i, 1], x[1]), Qmul(H[i, 2], x[2]), Qmul(H[i, 3], x[3]),
                                  Qmul(H[i, 4], x[4]), Qmul(H[i, 5], x[5]), Qmul(H[i, 6], x[6]), Qmul(H[i, 7], x[7]));
            out_s vr = v.real, vi = v.imag;
            if (vr.data.data != yb[i].real.data.data || vi.data.data != yb[i].imag.data.data) bad_b++;
        }
        std::printf("complex Qgemul TFComplexMul: %s\n", bad_tf ? "MISMATCH" : "OK");
        std::printf("complex Qgemul BasicComplexMul: %s\n", bad_b ? "MISMATCH" : "OK");
        bad += bad_tf + bad_b;

        // ---- mixed real / complex Qgemul: real 8x8 matrix times complex vector, and complex matrix times real vector ----
        {
            using R_t = Qu<dim<8, 8>, t>; using rv_t = Qu<dim<8>, t>;
            using pr = Qu<intBits<6>, fracBits<12>, QuMode<RND::CONV>>; using pi = Qu<intBits<6>, fracBits<11>, QuMode<RND::ZERO>, OfMode<SAT::ZERO>>;
            R_t Rm; rv_t rv; y_t ya, yb2;
            for (std::size_t i = 0; i < 64; i++) Rm[i].fill();
            for (std::size_t i = 0; i < 8; i++) rv[i].fill();
            Qgemul<QgemulAddArgs<clv>, QgemulMulArgs<realT<pr>, imagT<pi>>>(ya, Rm, x);       // real A, complex B
            Qgemul<QgemulMulArgs<pr>, QgemulAddArgs<clv>>(yb2, H, rv);                        // complex A, real B, one type for both parts
            int bad_rc = 0;
            for (int i = 0; i < 8; i++) {
                auto w = Qreduce<clv>(Qmul<realT<pr>, imagT<pi>>(Rm[i, 0], x[0]), Qmul<realT<pr>, imagT<pi>>(Rm[i, 1], x[1]), Qmul<realT<pr>, imagT<pi>>(Rm[i, 2], x[2]),
                                      Qmul<realT<pr>, imagT<pi>>(Rm[i, 3], x[3]), Qmul<realT<pr>, imagT<pi>>(Rm[i, 4], x[4]), Qmul<realT<pr>, imagT<pi>>(Rm[i, 5], x[5]),
                                      Qmul<realT<pr>, imagT<pi>>(Rm[i, 6], x[6]), Qmul<realT<pr>, imagT<pi>>(Rm[i, 7], x[7]));
                out_s wr = w.real, wi = w.imag;
                if (wr.data.data != ya[i].real.data.data || wi.data.data != ya[i].imag.data.data) bad_rc++;
                auto v = Qreduce<clv>(Qmul<pr>(H[i, 0], rv[0]), Qmul<pr>(H[i, 1], rv[1]), Qmul<pr>(H[i, 2], rv[2]), Qmul<pr>(H[i, 3], rv[3]),
                                      Qmul<pr>(H[i, 4], rv[4]), Qmul<pr>(H[i, 5], rv[5]), Qmul<pr>(H[i, 6], rv[6]), Qmul<pr>(H[i, 7], rv[7]));
                out_s vr = v.real, vi = v.imag;
                if (vr.data.data != yb2[i].real.data.data || vi.data.data != yb2[i].imag.data.data) bad_rc++;
            }
            auto ez = Qmul<realT<pr>, imagT<pi>>(rv, x);                                      // element-wise real (x) complex tensors
            auto ez2 = Qmul<pi, pr>(x, rv);
            for (std::size_t i = 0; i < 8; i++) {
                auto e1 = Qmul<realT<pr>, imagT<pi>>(rv[i], x[i]); auto e2 = Qmul<pi, pr>(x[i], rv[i]);
                if (e1.real.data.data != ez[i].real.data.data || e1.imag.data.data != ez[i].imag.data.data) bad_rc++;
                if (e2.real.data.data != ez2[i].real.data.data || e2.imag.data.data != ez2[i].imag.data.data) bad_rc++;
            }
            std::printf("real-by-complex Qgemul / Qmul: %s\n", bad_rc ? "MISMATCH" : "OK");
            bad += bad_rc;
        }

        // ---- element-wise tensor ops (the reference's lazy expressions, evaluated eagerly on the GPU) ----
        {
            using v_t = Qu<dim<100>, in_t>;
            v_t u, w; u.fill(); w.fill();
            auto s = Qadd<lv<0>>(u, w); auto d = Qsub<>(u, w); auto m = Qmul<mul_t>(u, w);
            int bad_ew = 0;
            for (std::size_t i = 0; i < 100; i++) {
                if (s[i].data.data != Qadd<lv<0>>(u[i], w[i]).data.data) bad_ew++;
                if (d[i].data.data != Qsub<>(u[i], w[i]).data.data) bad_ew++;
                if (m[i].data.data != Qmul<mul_t>(u[i], w[i]).data.data) bad_ew++;
            }
            std::printf("element-wise Qadd/Qsub/Qmul: %s\n", bad_ew ? "MISMATCH" : "OK");
            bad += bad_ew;
        }

        // ---- a vector-form Qreduce on a heap-backed (> 1000 elements) tensor: the reference cannot instantiate this ----
        using big_t = Qu<dim<4096>, Qu<intBits<8>, fracBits<8>>>;
        using rl = TypeList<Qu<intBits<9>, fracBits<8>>, Qu<intBits<10>, fracBits<7>, QuMode<RND::CONV>>, Qu<intBits<20>, fracBits<6>, QuMode<RND::INF>>>;
        big_t big; big.fill();
        auto r = Qreduce<rl>(big);
        std::printf("Qreduce 4096: raw %lld\n", (long long)r.data.data);
    } catch (const std::runtime_error &e) {
        std::printf("runtime_error: %s\n", e.what());
        return 3;
    }
    return bad ? 1 : 0;
}
