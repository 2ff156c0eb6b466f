// qb_fast_cgemv.cuh — compile-time-specialised batched complex mat-vec with TFComplexMul products
// (BASELINE config 4: 256x256 complex H times x, 3 multipliers / 5 adders per product,
// QuBLAS.h:3522-3534, accumulated by the component-wise adjacent-pair tree, QuBLAS.h:3549-3563).
//
//   * one CTA per instance, one thread per output row i; H is column-major AoS {re, im}, so for a
//     fixed k the CTA reads M consecutive 8-byte pairs: fully coalesced, 8 independent loads in
//     flight per thread;
//   * the x-only part of the TF dataflow (c + d, Qadd<cdT>) is computed once per instance into
//     shared memory and broadcast;
//   * 8 complex products per chunk go through a fully unrolled 3-level tree, chunks through the
//     level-indexed register stack (two stacks: real, imaginary); every quantizer is an immediate.
// This kernel sits near the balance point of HBM (8 B per complex MAC) and the integer pipes
// (~25 integer ops per complex MAC): see DESIGN.md.
#pragma once
#include "qb_device.cuh"

namespace qb {

#ifndef QB_FADD
#define QB_FADD 1
#endif
// x + y (or x - y) of two exact floats below 2^23 with the stage's SAT::TCPL clamp: exact, on the FP pipe
QB_DEV float fadd_stage(float x, float y, const QStage &q, bool sub)
{
    float r = sub ? x - y : x + y;
    if (q.need_hi) r = fminf(r, (float)q.hi);
    if (q.need_lo) r = fmaxf(r, (float)q.lo);
    return r;
}

template <class Cfg, int L, int NTOP, typename I>
QB_DEV void cstack_push(I (&sr)[NTOP > 0 ? NTOP : 1], I (&si)[NTOP > 0 ? NTOP : 1], uint32_t c, I vr, I vi, I &rr, I &ri)
{
    if constexpr (L >= NTOP) {
        rr = vr; ri = vi;
    } else {
        if ((c >> L) & 1) {
            constexpr AddStage lr = Cfg::lvl_re(3 + L), li = Cfg::lvl_im(3 + L);
            I mr = merge_apply<I, true>(sr[L], vr, lr);
            I mi = merge_apply<I, true>(si[L], vi, li);
            cstack_push<Cfg, L + 1, NTOP, I>(sr, si, c, mr, mi, rr, ri);
        } else {
            sr[L] = vr; si[L] = vi;
        }
    }
}

template <class Cfg>
QB_DEV void cgemv_rows_body(const int2 *__restrict__ H, const int2 *__restrict__ X, int2 *__restrict__ Y, uint64_t batch)
{
    typedef typename Cfg::I I;
    constexpr int M = Cfg::M, K = Cfg::K, NTOP = Cfg::NLEV - 3;
    static_assert(Cfg::N == 1 && (K & (K - 1)) == 0 && K >= 8, "mat-vec with power-of-two K");
    __shared__ int xc[K], xd[K], xcd[K];          // c, d, c+d: raw ints, or exact floats where the product runs on the FMA pipe
    constexpr TfProg tf = Cfg::tf();
    constexpr AddStage r0 = Cfg::lvl_re(0), r1 = Cfg::lvl_re(1), r2 = Cfg::lvl_re(2);
    constexpr AddStage i0 = Cfg::lvl_im(0), i1 = Cfg::lvl_im(1), i2 = Cfg::lvl_im(2);
    constexpr QStage ore = Cfg::out_re(), oim = Cfg::out_im();
    const int i = threadIdx.x;

    for (uint64_t inst = blockIdx.x; inst < batch; inst += gridDim.x) {
        __syncthreads();                                   // previous instance finished with xc/xd/xcd
        for (int k = i; k < K; k += M) {
            const int2 x = __ldg(X + inst * (uint64_t)Cfg::SB + k);          // batch strides in complex elements; 0 = shared operand
            const int cdv = (int)add_apply<I, true>((I)x.x, (I)x.y, tf.cd);       // Qadd<cdT>(c, d), QuBLAS.h:3525
            xc[k] = Cfg::FP_A ? __float_as_int(__int2float_rn(x.x)) : x.x;
            xd[k] = Cfg::FP_C ? __float_as_int(__int2float_rn(x.y)) : x.y;
            xcd[k] = Cfg::FP_B ? __float_as_int(__int2float_rn(cdv)) : cdv;
        }
        __syncthreads();
        const int2 *Hi = H + inst * (uint64_t)Cfg::SA + i;
        I sr[NTOP > 0 ? NTOP : 1], si[NTOP > 0 ? NTOP : 1];
        I rr = 0, ri = 0;
        // chunk loop unrolled by 4: the two lowest stack levels fold at compile time and 32 loads are in flight per thread
        // (config 4, NVRTC build: unroll 1 / 2 / 4 = 770 / 780 / 796 G complex MAC/s)
#ifndef QB_CGEMV_UNROLL
#define QB_CGEMV_UNROLL 4
#endif
        constexpr int CHUNK_UNROLL = QB_CGEMV_UNROLL;
#pragma unroll CHUNK_UNROLL
        for (uint32_t c = 0; c < (uint32_t)(K / 8); ++c) {
            int2 h[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) h[u] = __ldg(Hi + (uint64_t)(c * 8 + u) * M);
            I pr[8], pi[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int k = c * 8 + u;
                const I a = h[u].x, b = h[u].y, cc = xc[k], d = xd[k], cd = xcd[k];
                const float af = __int2float_rn((int)a), bf = __int2float_rn((int)b);     // exact; dead code unless an FP path uses them
                float abf, baf;
                if constexpr (Cfg::FADD_AB && QB_FADD) abf = fadd_stage(af, bf, tf.ab.q, false);          // :3524 on the FP pipe
                else abf = __int2float_rn((int)add_apply<I, true>(a, b, tf.ab));
                if constexpr (Cfg::FADD_BA && QB_FADD) baf = fadd_stage(bf, af, tf.ba.q, true);           // :3526 (default merger: baT is ignored upstream)
                else baf = __int2float_rn((int)add_apply<I, true>(b, a, tf.ba));
                I A, Bv, Cv;
                if constexpr (Cfg::FP_A) A = (I)(fp_mul_apply<Cfg::SH_A, true>(abf, __int_as_float((int)cc), tf.A) - fp_kbits<Cfg::SH_A>());
                else A = mul_apply<I, true>(add_apply<I, true>(a, b, tf.ab), cc, tf.A);
                if constexpr (Cfg::FP_B) Bv = (I)(fp_mul_apply<Cfg::SH_B, true>(__int_as_float((int)cd), bf, tf.B) - fp_kbits<Cfg::SH_B>());
                else Bv = mul_apply<I, true>(cd, b, tf.B);
                if constexpr (Cfg::FP_C) Cv = (I)(fp_mul_apply<Cfg::SH_C, true>(baf, __int_as_float((int)d), tf.C) - fp_kbits<Cfg::SH_C>());
                else Cv = mul_apply<I, true>(add_apply<I, true>(b, a, tf.ba), d, tf.C);
                pr[u] = add_apply<I, true>(A, Bv, tf.re);                // :3528
                pi[u] = add_apply<I, true>(Bv, Cv, tf.im);               // :3529
            }
            I vr = merge_apply<I, true>(
                merge_apply<I, true>(merge_apply<I, true>(pr[0], pr[1], r0), merge_apply<I, true>(pr[2], pr[3], r0), r1),
                merge_apply<I, true>(merge_apply<I, true>(pr[4], pr[5], r0), merge_apply<I, true>(pr[6], pr[7], r0), r1), r2);
            I vi = merge_apply<I, true>(
                merge_apply<I, true>(merge_apply<I, true>(pi[0], pi[1], i0), merge_apply<I, true>(pi[2], pi[3], i0), i1),
                merge_apply<I, true>(merge_apply<I, true>(pi[4], pi[5], i0), merge_apply<I, true>(pi[6], pi[7], i0), i1), i2);
            cstack_push<Cfg, 0, NTOP, I>(sr, si, c, vr, vi, rr, ri);
        }
        int2 o;
        o.x = (int)conv_apply<I, true>(rr, ore);
        o.y = (int)conv_apply<I, true>(ri, oim);
        Y[inst * (uint64_t)Cfg::SC + i] = o;
    }
}

// Occupancy target: 1280 resident threads per SM (5 CTAs of 256 rows), i.e. at most 48 registers per thread.  The
// kernel sits at the HBM / ALU balance point and the extra CTA of loads in flight is worth 7 % (config 4: 716 -> 773
// G complex MAC/s, 0.88 -> 0.95 of the measured HBM bandwidth); found because NVRTC happened to allocate 47 registers
// where nvcc took 52.
#ifndef QB_CGEMV_THREADS_PER_SM
#define QB_CGEMV_THREADS_PER_SM 1280
#endif
constexpr int cgemv_min_blocks(int rows)
{
    return QB_CGEMV_THREADS_PER_SM / rows < 1 ? 1 : (QB_CGEMV_THREADS_PER_SM / rows > 16 ? 16 : QB_CGEMV_THREADS_PER_SM / rows);
}

template <class Cfg>
__global__ void __launch_bounds__(Cfg::M, cgemv_min_blocks(Cfg::M))
k_cgemv_rows(const int2 *__restrict__ H, const int2 *__restrict__ X, int2 *__restrict__ Y, uint64_t batch)
{
    cgemv_rows_body<Cfg>(H, X, Y, batch);
}

#ifndef QB_JIT
template <class Cfg>
static int launch_cgemv_rows(const void *A, const void *B, void *C, uint64_t batch, cudaStream_t s)
{
    static int grid_caps[64] = { 0 };                 // per device: function attributes and occupancy belong to a context
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    int &grid_cap = grid_caps[cur_dev & 63];
    if (!grid_cap) {
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cgemv_rows<Cfg>, Cfg::M, 0);
        if (e != cudaSuccess) return (int)e;
        grid_cap = sms * (per_sm > 0 ? per_sm : 1);
    }
    const unsigned grid = (unsigned)(batch < (uint64_t)grid_cap ? batch : (uint64_t)grid_cap);
    k_cgemv_rows<Cfg><<<grid, Cfg::M, 0, s>>>((const int2 *)A, (const int2 *)B, (int2 *)C, batch);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) count_launch(1);
    return (int)e;
}

#endif  // QB_JIT

}  // namespace qb
