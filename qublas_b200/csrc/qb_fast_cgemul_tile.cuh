// qb_fast_cgemul_tile.cuh — compile-time-specialised batched COMPLEX Qgemul with N > 1 on register tiles
// (TFComplexMul QuBLAS.h:3510-3535 or BasicComplexMul :3426-3446 products, component-wise adjacent-pair trees :3549-3568).
//
//   * one CTA per instance, persistent over the batch; the A and B instances (column-major AoS {re, im}, 8 bytes per
//     element) are staged by 1-D TMA bulk copies into a two-stage shared-memory ring, so the next instance streams in
//     while the current one is computed;
//   * each thread owns a 4 x 2 tile of C: per pair of k it reads four rows of A (two 16-byte loads per k) and two
//     columns of B (one 16-byte load per column holds both k), forms the operand-only nodes of the product once per
//     operand element (a+b and b-a per A element, c+d per B element: shared by the tile's columns resp. rows) and
//     runs the three (or four) quantized products and the two quantized sums per complex MAC;
//   * leaves go through two level-indexed register stacks (real, imaginary): the three lowest levels are indexed by the
//     position inside an 8-leaf chunk (compile-time), the levels above by the chunk counter (chunk loop unrolled);
//   * lanes are mapped to tiles so that four consecutive lanes read at most two distinct 16-byte chunks of either
//     operand (the rule measured in tools/lds_probe.cu).
// Every shift / bound / mode is an immediate; same device functions as the thread-per-output kernel it replaces for
// these shapes, so the two are compared bit for bit (QB_F_NO_SPECIALIZE) besides the oracle.
#pragma once
#include "qb_fast_gemul.cuh"

namespace qb {

template <class Cfg>
struct CGemulTile {
    static constexpr int M = (int)Cfg::M, N = (int)Cfg::N, K = (int)Cfg::K;
    static constexpr int TM = 4, TN = 2;
    static constexpr int TPI = (M / TM) * (N / TN);
    static constexpr int BLOCK = TPI;
    static constexpr int NCH = K / 8, NTOP = Cfg::NLEV - 3;
    static constexpr uint32_t A_BYTES = (uint32_t)M * K * 8, B_BYTES = (uint32_t)K * N * 8, STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr uint32_t SMEM_BYTES = 2 * STAGE_BYTES + 64;
    static constexpr bool PATCH = (M / TM) % 8 == 0 && (N / TN) % 4 == 0;
    static_assert(M % TM == 0 && N % TN == 0 && K % 8 == 0 && (K & (K - 1)) == 0 && BLOCK % 32 == 0 && BLOCK <= 1024, "tile shape");
    static_assert(Cfg::KIND == 0 || Cfg::KIND == 1, "complex (x) complex products");
};

// stack over the three lowest levels, indexed by the leaf position u inside the chunk (compile-time after unrolling)
template <class Cfg, int L, typename I>
QB_DEV void clow_push(I (&sr)[3], I (&si)[3], int u, I vr, I vi, I &outr, I &outi)
{
    if constexpr (L >= 3) {
        outr = vr; outi = vi;
    } else {
        if ((u >> L) & 1) {
            constexpr AddStage lr = Cfg::lvl_re(L), li = Cfg::lvl_im(L);
            clow_push<Cfg, L + 1, I>(sr, si, u, merge_apply<I, true>(sr[L], vr, lr), merge_apply<I, true>(si[L], vi, li), outr, outi);
        } else {
            sr[L] = vr; si[L] = vi;
        }
    }
}
// stack over the levels above a chunk, indexed by the chunk counter
template <class Cfg, int L, int NTOP, typename I>
QB_DEV void ctop_push(I (&sr)[NTOP > 0 ? NTOP : 1], I (&si)[NTOP > 0 ? NTOP : 1], uint32_t c, I vr, I vi, I &rr, I &ri)
{
    if constexpr (L >= NTOP) {
        rr = vr; ri = vi;
    } else {
        if ((c >> L) & 1) {
            constexpr AddStage lr = Cfg::lvl_re(3 + L), li = Cfg::lvl_im(3 + L);
            ctop_push<Cfg, L + 1, NTOP, I>(sr, si, c, merge_apply<I, true>(sr[L], vr, lr), merge_apply<I, true>(si[L], vi, li), rr, ri);
        } else {
            sr[L] = vr; si[L] = vi;
        }
    }
}

template <class Cfg>
QB_DEV void cgemul_tile_body(const int2 *__restrict__ A, const int2 *__restrict__ B, int2 *__restrict__ C, uint64_t batch)
{
    typedef CGemulTile<Cfg> T;
    typedef typename Cfg::I I;
    constexpr int M = T::M, K = T::K, NTOP = T::NTOP;
    constexpr TfProg tf = Cfg::tf();
    constexpr BcmProg bcm = Cfg::bcm();
    constexpr QStage ore = Cfg::out_re(), oim = Cfg::out_im();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bars = (uint64_t *)(smem_raw + 2 * T::STAGE_BYTES);

    const int tid = threadIdx.x;
    int tx, ty;
    if constexpr (T::PATCH) {
        constexpr int WX = (M / 4) / 8;
        const int wp = tid >> 5;
        tx = (wp % WX) * 8 + ((tid & 1) | ((tid >> 1) & 2) | ((tid >> 2) & 4));
        ty = (wp / WX) * 4 + (((tid >> 1) & 1) | ((tid >> 2) & 2));
    } else {
        tx = tid % (M / 4); ty = tid / (M / 4);
    }
    const int i0 = tx * 4, j0 = ty * 2;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](uint64_t inst, int s) {
        unsigned char *dst = smem_raw + (size_t)s * T::STAGE_BYTES;
        mbar_expect_tx(&bars[s], T::STAGE_BYTES);
        tma_load_1d(dst, A + inst * (uint64_t)Cfg::SA, T::A_BYTES, &bars[s]);
        tma_load_1d(dst + T::A_BYTES, B + inst * (uint64_t)Cfg::SB, T::B_BYTES, &bars[s]);
    };
    uint64_t inst = blockIdx.x;
    if (inst < batch && tid == 0) issue(inst, 0);

    for (uint32_t it = 0; inst < batch; inst += gridDim.x, ++it) {
        const int s = it & 1;
        const uint64_t next = inst + gridDim.x;
        if (tid == 0 && next < batch) issue(next, s ^ 1);          // stage s^1 was released by the barrier that ended the previous iteration
        mbar_wait(&bars[s], (it >> 1) & 1);
        const int4 *As4 = (const int4 *)(smem_raw + (size_t)s * T::STAGE_BYTES);          // two complex elements per 16 bytes
        const int4 *Bs4 = (const int4 *)(smem_raw + (size_t)s * T::STAGE_BYTES + T::A_BYTES);

        I tr[4][2][NTOP > 0 ? NTOP : 1], ti[4][2][NTOP > 0 ? NTOP : 1];
        I rr[4][2], ri[4][2];
        constexpr int CHUNK_UNROLL = T::NCH <= 8 ? T::NCH : 1;
#pragma unroll CHUNK_UNROLL
        for (uint32_t c = 0; c < (uint32_t)T::NCH; ++c) {
            I lr[4][2][3], li[4][2][3];
#pragma unroll
            for (int up = 0; up < 8; up += 2) {
                const int k0 = (int)c * 8 + up;
                // A: rows i0 .. i0+3 at k0 and k0+1 (column-major: rows are contiguous); B: k0, k0+1 of columns j0, j0+1
                int4 aq[2][2], bq[2];
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) { aq[kk][0] = As4[((k0 + kk) * M + i0) >> 1]; aq[kk][1] = As4[(((k0 + kk) * M + i0) >> 1) + 1]; }
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) bq[jj] = Bs4[((j0 + jj) * K + k0) >> 1];
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    const int u = up + kk;
                    // operand-only nodes: once per operand element
                    I a[4], b[4], ab[4], ba[4];
#pragma unroll
                    for (int ii = 0; ii < 4; ++ii) {
                        a[ii] = (ii & 1) ? aq[kk][ii >> 1].z : aq[kk][ii >> 1].x;
                        b[ii] = (ii & 1) ? aq[kk][ii >> 1].w : aq[kk][ii >> 1].y;
                        if constexpr (Cfg::KIND == 0) {
                            ab[ii] = add_apply<I, true>(a[ii], b[ii], tf.ab);            // QuBLAS.h:3524 Qadd<abT>(a, b)
                            ba[ii] = add_apply<I, true>(b[ii], a[ii], tf.ba);            // :3526 Qsub(b, a)
                        }
                    }
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) {
                        const I cc = kk ? bq[jj].z : bq[jj].x, d = kk ? bq[jj].w : bq[jj].y;
                        I cd = 0;
                        if constexpr (Cfg::KIND == 0) cd = add_apply<I, true>(cc, d, tf.cd);          // :3525 Qadd<cdT>(c, d)
#pragma unroll
                        for (int ii = 0; ii < 4; ++ii) {
                            I vr, vi;
                            if constexpr (Cfg::KIND == 0) {
                                const I Av = mul_apply<I, true>(ab[ii], cc, tf.A);
                                const I Bv = mul_apply<I, true>(cd, b[ii], tf.B);
                                const I Cv = mul_apply<I, true>(ba[ii], d, tf.C);
                                vr = add_apply<I, true>(Av, Bv, tf.re);                  // :3528
                                vi = add_apply<I, true>(Bv, Cv, tf.im);                  // :3529
                            } else {
                                const I ac = mul_apply<I, true>(a[ii], cc, bcm.ac), bd = mul_apply<I, true>(b[ii], d, bcm.bd);
                                const I ad = mul_apply<I, true>(a[ii], d, bcm.ad), bc = mul_apply<I, true>(b[ii], cc, bcm.bc);
                                vr = add_apply<I, true>(ac, bd, bcm.re);                 // :3439
                                vi = add_apply<I, true>(ad, bc, bcm.im);                 // :3440
                            }
                            I nr, ni;
                            clow_push<Cfg, 0, I>(lr[ii][jj], li[ii][jj], u, vr, vi, nr, ni);
                            if (u == 7) ctop_push<Cfg, 0, NTOP, I>(tr[ii][jj], ti[ii][jj], c, nr, ni, rr[ii][jj], ri[ii][jj]);
                        }
                    }
                }
            }
        }
        // converting assignment into C's element types; rows i0, i0+1 and i0+2, i0+3 of a column are 16 bytes each
        int2 *Ci = C + inst * (uint64_t)Cfg::SC;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                int4 o;
                o.x = (int)conv_apply<I, true>(rr[2 * h][jj], ore); o.y = (int)conv_apply<I, true>(ri[2 * h][jj], oim);
                o.z = (int)conv_apply<I, true>(rr[2 * h + 1][jj], ore); o.w = (int)conv_apply<I, true>(ri[2 * h + 1][jj], oim);
                *(int4 *)(Ci + (size_t)(j0 + jj) * M + i0 + 2 * h) = o;
            }
        }
        __syncthreads();      // everyone is done reading stage s before it is refilled
    }
}

// ---- TFComplexMul with all three products on the FP32 FMA pipe ------------------------------------------------------
// When Qmul<abcT>, Qmul<badT> and Qmul<cdbT> are exact as one FFMA with a magic constant (qb_codegen.cpp:fp_mul_ok, the
// same proof as for the mat-vec kernel), the operand-only nodes move out of the inner loop altogether: one pass per staged
// instance turns A into three float planes (a+b, b, b-a -- the real part a alone is never multiplied) and B into
// (c, c+d, d), each node computed by its integer quantizer and converted exactly.  The inner loop is then 3 FFMA, the two
// quantized sums and the two tree merges per complex MAC.  B planes are K-contiguous per column and XOR-swizzled by the
// column block so that the four column blocks a warp reads fall into different banks.
template <class Cfg>
struct CGemulTileFp {
    typedef CGemulTile<Cfg> T;
    static constexpr int M = T::M, N = T::N, K = T::K;
    static constexpr uint32_t RAW_BYTES = T::STAGE_BYTES;                                  // one staging buffer
    static constexpr uint32_t PLANE_A = (uint32_t)M * K * 4, PLANE_B = (uint32_t)K * N * 4;
    static constexpr int NP = Cfg::KIND == 0 ? 3 : 2;               // float planes per operand: TF (a+b, b, b-a | c, c+d, d), Basic (a, b | c, d)
    static constexpr uint32_t SMEM_BYTES = RAW_BYTES + NP * PLANE_A + NP * PLANE_B + 64;
    static_assert(K % 16 == 0, "swizzled B planes need at least four 16-byte chunks per column");
    static_assert((Cfg::KIND == 0 && Cfg::FP_A && Cfg::FP_B && Cfg::FP_C) || (Cfg::KIND == 1 && Cfg::FP_BCM), "every product on the FP pipe");
};

template <class Cfg>
QB_DEV void cgemul_tile_fp_body(const int2 *__restrict__ A, const int2 *__restrict__ B, int2 *__restrict__ C, uint64_t batch)
{
    typedef CGemulTile<Cfg> T;
    typedef CGemulTileFp<Cfg> F;
    typedef typename Cfg::I I;
    constexpr int M = T::M, N = T::N, K = T::K, NTOP = T::NTOP;
    constexpr TfProg tf = Cfg::tf();
    constexpr QStage ore = Cfg::out_re(), oim = Cfg::out_im();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr bool TF = Cfg::KIND == 0;
    constexpr BcmProg bcm = Cfg::bcm();
    // A planes [k][i]: TF (a+b, b, b-a), Basic (a, b, -);  B planes [j][k], swizzled: TF (c, c+d, d), Basic (c, -, d)
    float *pab = (float *)(smem_raw + F::RAW_BYTES), *pb = pab + M * K, *pba = TF ? pb + M * K : pb;
    float *pc = (float *)(smem_raw + F::RAW_BYTES + F::NP * F::PLANE_A), *pcd = TF ? pc + K * N : pc, *pd = TF ? pcd + K * N : pc + K * N;
    uint64_t *bar = (uint64_t *)(smem_raw + F::RAW_BYTES + F::NP * F::PLANE_A + F::NP * F::PLANE_B);

    const int tid = threadIdx.x;
    int tx, ty;
    if constexpr (T::PATCH) {
        constexpr int WX = (M / 4) / 8;
        const int wp = tid >> 5;
        tx = (wp % WX) * 8 + ((tid & 1) | ((tid >> 1) & 2) | ((tid >> 2) & 4));
        ty = (wp / WX) * 4 + (((tid >> 1) & 1) | ((tid >> 2) & 2));
    } else {
        tx = tid % (M / 4); ty = tid / (M / 4);
    }
    const int i0 = tx * 4, j0 = ty * 2;

    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](uint64_t inst) {
        mbar_expect_tx(bar, T::STAGE_BYTES);
        tma_load_1d(smem_raw, A + inst * (uint64_t)Cfg::SA, T::A_BYTES, bar);
        tma_load_1d(smem_raw + T::A_BYTES, B + inst * (uint64_t)Cfg::SB, T::B_BYTES, bar);
    };
    uint64_t inst = blockIdx.x;
    if (inst < batch && tid == 0) issue(inst);

    // per-thread B chunk offsets (in float4 units): chunk (kc % 4) ^ (column block % 4) within an aligned group of four
    const int swb = ty & 3;
    const float4 *bcol[3][2];
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) { bcol[0][jj] = (const float4 *)(pc + (j0 + jj) * K); bcol[1][jj] = (const float4 *)(pcd + (j0 + jj) * K); bcol[2][jj] = (const float4 *)(pd + (j0 + jj) * K); }

    for (uint32_t it = 0; inst < batch; inst += gridDim.x, ++it) {
        mbar_wait(bar, it & 1);
        // ---- pass: operand-only nodes as exact floats ----
        {
            const int4 *rawA = (const int4 *)smem_raw;                                      // two complex elements per 16 bytes
            for (int e = tid; e < M * K / 4; e += T::BLOCK) {                              // four consecutive rows of one k
                const int4 q0 = rawA[2 * e], q1 = rawA[2 * e + 1];
                const I a[4] = { q0.x, q0.z, q1.x, q1.z }, b[4] = { q0.y, q0.w, q1.y, q1.w };
                float fab[4], fb[4], fba[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    fab[u] = __int2float_rn(TF ? (int)add_apply<I, true>(a[u], b[u], tf.ab) : (int)a[u]);    // QuBLAS.h:3524 Qadd<abT>(a, b)
                    fb[u] = __int2float_rn((int)b[u]);
                    if constexpr (TF) fba[u] = __int2float_rn((int)add_apply<I, true>(b[u], a[u], tf.ba));    // :3526 Qsub(b, a)
                }
                ((float4 *)pab)[e] = make_float4(fab[0], fab[1], fab[2], fab[3]);
                ((float4 *)pb)[e] = make_float4(fb[0], fb[1], fb[2], fb[3]);
                if constexpr (TF) ((float4 *)pba)[e] = make_float4(fba[0], fba[1], fba[2], fba[3]);
            }
            const int4 *rawB = (const int4 *)(smem_raw + T::A_BYTES);
            for (int e = tid; e < K * N / 4; e += T::BLOCK) {                              // four consecutive k of one column
                const int4 q0 = rawB[2 * e], q1 = rawB[2 * e + 1];
                const I c[4] = { q0.x, q0.z, q1.x, q1.z }, d[4] = { q0.y, q0.w, q1.y, q1.w };
                float fc[4], fcd[4], fd[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    fc[u] = __int2float_rn((int)c[u]);
                    if constexpr (TF) fcd[u] = __int2float_rn((int)add_apply<I, true>(c[u], d[u], tf.cd));    // :3525 Qadd<cdT>(c, d)
                    fd[u] = __int2float_rn((int)d[u]);
                }
                const int j = e / (K / 4), ch = e % (K / 4);
                const int o = j * (K / 4) + (ch ^ ((j >> 1) & 3));                         // swizzle by the column block (two columns per tile)
                ((float4 *)pc)[o] = make_float4(fc[0], fc[1], fc[2], fc[3]);
                if constexpr (TF) ((float4 *)pcd)[o] = make_float4(fcd[0], fcd[1], fcd[2], fcd[3]);
                ((float4 *)pd)[o] = make_float4(fd[0], fd[1], fd[2], fd[3]);
            }
        }
        __syncthreads();
        const uint64_t next = inst + gridDim.x;
        if (tid == 0 && next < batch) issue(next);                 // the pass has consumed the staged instance

        I tr[4][2][NTOP > 0 ? NTOP : 1], ti[4][2][NTOP > 0 ? NTOP : 1];
        I rr[4][2], ri[4][2];
        constexpr int CHUNK_UNROLL = T::NCH <= 8 ? T::NCH : 1;
#pragma unroll CHUNK_UNROLL
        for (uint32_t c = 0; c < (uint32_t)T::NCH; ++c) {
            I lr[4][2][3], li[4][2][3];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int kc = 2 * (int)c + h;                     // 16-byte chunk of k
                float4 fcq[2], fcdq[2], fdq[2];
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const int o = (kc & ~3) | ((kc & 3) ^ swb);
                    fcq[jj] = bcol[0][jj][o]; fdq[jj] = bcol[2][jj][o];
                    if constexpr (TF) fcdq[jj] = bcol[1][jj][o];
                }
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const int k = 4 * kc + kk, u = 4 * h + kk;
                    const float4 fab = *(const float4 *)(pab + k * M + i0), fb = *(const float4 *)(pb + k * M + i0);
                    float4 fba = fb;
                    if constexpr (TF) fba = *(const float4 *)(pba + k * M + i0);
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) {
                        const float cc = (&fcq[jj].x)[kk], d = (&fdq[jj].x)[kk];
                        float cd = cc;
                        if constexpr (TF) cd = (&fcdq[jj].x)[kk];
#pragma unroll
                        for (int ii = 0; ii < 4; ++ii) {
                            I vr, vi;
                            if constexpr (TF) {
                                const I Av = (I)(fp_mul_apply<Cfg::SH_A, true>((&fab.x)[ii], cc, tf.A) - fp_kbits<Cfg::SH_A>());
                                const I Bv = (I)(fp_mul_apply<Cfg::SH_B, true>(cd, (&fb.x)[ii], tf.B) - fp_kbits<Cfg::SH_B>());
                                const I Cv = (I)(fp_mul_apply<Cfg::SH_C, true>((&fba.x)[ii], d, tf.C) - fp_kbits<Cfg::SH_C>());
                                vr = add_apply<I, true>(Av, Bv, tf.re);                      // :3528
                                vi = add_apply<I, true>(Bv, Cv, tf.im);                      // :3529
                            } else {                                                         // fab = a, fb = b (BasicComplexMul, :3437-3445)
                                const I ac = (I)(fp_mul_apply<Cfg::SH_AC, true>((&fab.x)[ii], cc, bcm.ac) - fp_kbits<Cfg::SH_AC>());
                                const I bd = (I)(fp_mul_apply<Cfg::SH_BD, true>((&fb.x)[ii], d, bcm.bd) - fp_kbits<Cfg::SH_BD>());
                                const I ad = (I)(fp_mul_apply<Cfg::SH_AD, true>((&fab.x)[ii], d, bcm.ad) - fp_kbits<Cfg::SH_AD>());
                                const I bc = (I)(fp_mul_apply<Cfg::SH_BC, true>((&fb.x)[ii], cc, bcm.bc) - fp_kbits<Cfg::SH_BC>());
                                vr = add_apply<I, true>(ac, bd, bcm.re);                     // :3439
                                vi = add_apply<I, true>(ad, bc, bcm.im);                     // :3440
                            }
                            I nr, ni;
                            clow_push<Cfg, 0, I>(lr[ii][jj], li[ii][jj], u, vr, vi, nr, ni);
                            if (u == 7) ctop_push<Cfg, 0, NTOP, I>(tr[ii][jj], ti[ii][jj], c, nr, ni, rr[ii][jj], ri[ii][jj]);
                        }
                    }
                }
            }
        }
        int2 *Ci = C + inst * (uint64_t)Cfg::SC;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                int4 o;
                o.x = (int)conv_apply<I, true>(rr[2 * h][jj], ore); o.y = (int)conv_apply<I, true>(ri[2 * h][jj], oim);
                o.z = (int)conv_apply<I, true>(rr[2 * h + 1][jj], ore); o.w = (int)conv_apply<I, true>(ri[2 * h + 1][jj], oim);
                *(int4 *)(Ci + (size_t)(j0 + jj) * M + i0 + 2 * h) = o;
            }
        }
        __syncthreads();      // the planes are free for the next pass
    }
}

}  // namespace qb
