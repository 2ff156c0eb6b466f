// qb_fast_ew.cuh — compile-time-specialised element-wise kernels: the complex products on tensors (TFComplexMul
// QuBLAS.h:3522-3534, BasicComplexMul :3437-3445, real (x) complex :3600-3637; tensor form :3783-3811), element-wise
// Qmul / Qadd / Qsub of real tensors (:3741-3811, :4080-4156) and the converting assignment (:2398-2411, :2759-2830).
// These are streaming kernels: with every quantizer an immediate the arithmetic (8 stages for a TF product) fits under
// the HBM time of the 24 bytes it moves, which the runtime-descriptor kernels do not manage.
// Each thread keeps four independent elements in flight; streaming loads and stores (every byte is touched once).
#pragma once
#include "qb_device.cuh"

namespace qb {

#ifndef QB_EW_U
#define QB_EW_U 4
#endif

template <typename W> QB_DEV void ew_load_pair(const W *p, uint64_t e, W &re, W &im) { re = __ldcs(p + 2 * e); im = __ldcs(p + 2 * e + 1); }
template <> QB_DEV void ew_load_pair<int32_t>(const int32_t *p, uint64_t e, int32_t &re, int32_t &im)
{
    const int2 v = __ldcs((const int2 *)p + e);
    re = v.x; im = v.y;
}
template <typename W> QB_DEV void ew_store_pair(W *p, uint64_t e, W re, W im) { __stcs(p + 2 * e, re); __stcs(p + 2 * e + 1, im); }
template <> QB_DEV void ew_store_pair<int32_t>(int32_t *p, uint64_t e, int32_t re, int32_t im) { __stcs((int2 *)p + e, make_int2(re, im)); }

// complex product per element; a real operand (KIND 2: x, KIND 3: y) holds one word per element
template <class Cfg>
QB_DEV void cew_body(const typename Cfg::WX *__restrict__ x, const typename Cfg::WY *__restrict__ y, typename Cfg::WO *__restrict__ out, uint64_t n)
{
    typedef typename Cfg::I I;
    typedef typename Cfg::WX WX; typedef typename Cfg::WY WY; typedef typename Cfg::WO WO;
    constexpr int KIND = Cfg::KIND, U = QB_EW_U;
    constexpr TfProg tf = Cfg::tf();
    constexpr BcmProg bcm = Cfg::bcm();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e0 < n; e0 += U * stride) {
        WX a[U], b[U]; WY c[U], d[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint64_t e = e0 + u * stride;
            a[u] = 0; b[u] = 0; c[u] = 0; d[u] = 0;
            if (e < n) {
                if constexpr (KIND == 2) a[u] = __ldcs(x + e); else ew_load_pair<WX>(x, e, a[u], b[u]);
                if constexpr (KIND == 3) c[u] = __ldcs(y + e); else ew_load_pair<WY>(y, e, c[u], d[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint64_t e = e0 + u * stride;
            if (e < n) {
                I re, im;
                cprod_apply<I, true>(KIND, (I)a[u], (I)b[u], (I)c[u], (I)d[u], tf, bcm, re, im);
                ew_store_pair<WO>(out, e, (WO)re, (WO)im);
            }
        }
    }
}

// out[i] = Qmul / Qadd / Qsub <to>(a[i], b[i]); Cfg::OP: 0 mul, 1 add, 2 sub (the sub flag lives in the stage)
template <class Cfg>
QB_DEV void ew_body(const typename Cfg::WX *__restrict__ x, const typename Cfg::WY *__restrict__ y, typename Cfg::WO *__restrict__ out, uint64_t n)
{
    typedef typename Cfg::I I;
    constexpr int U = QB_EW_U;
    constexpr AddStage st = Cfg::stage();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e0 < n; e0 += U * stride) {
        typename Cfg::WX a[U]; typename Cfg::WY b[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint64_t e = e0 + u * stride;
            a[u] = 0; b[u] = 0;
            if (e < n) { a[u] = __ldcs(x + e); b[u] = __ldcs(y + e); }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint64_t e = e0 + u * stride;
            if (e < n) {
                const I r = Cfg::OP == 0 ? mul_apply<I, true>((I)a[u], (I)b[u], st.q) : add_apply<I, true>((I)a[u], (I)b[u], st);
                __stcs(out + e, (typename Cfg::WO)r);
            }
        }
    }
}

// out[i] = To(in[i]): the converting assignment; `y` is unused (one signature for all element-wise entry points)
template <class Cfg>
QB_DEV void conv_body(const typename Cfg::WX *__restrict__ x, const typename Cfg::WY *__restrict__, typename Cfg::WO *__restrict__ out, uint64_t n)
{
    typedef typename Cfg::I I;
    constexpr int U = QB_EW_U;
    constexpr QStage q = Cfg::stage().q;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e0 < n; e0 += U * stride) {
        typename Cfg::WX a[U];
#pragma unroll
        for (int u = 0; u < U; u++) { const uint64_t e = e0 + u * stride; a[u] = e < n ? __ldcs(x + e) : (typename Cfg::WX)0; }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint64_t e = e0 + u * stride;
            if (e < n) __stcs(out + e, (typename Cfg::WO)conv_apply<I, true>((I)a[u], q));
        }
    }
}

// out[i] = ANUS::Qpoly / Qapprox (in[i]) (QuBLAS.h:4829-4897); `y` is unused
template <class Cfg>
QB_DEV void poly_body(const typename Cfg::WX *__restrict__ x, const typename Cfg::WY *__restrict__, typename Cfg::WO *__restrict__ out, uint64_t n)
{
    typedef typename Cfg::I I;
    constexpr int U = QB_EW_U;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e0 < n; e0 += U * stride) {
        typename Cfg::WX a[U];
#pragma unroll
        for (int u = 0; u < U; u++) { const uint64_t e = e0 + u * stride; a[u] = e < n ? __ldcs(x + e) : (typename Cfg::WX)0; }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint64_t e = e0 + u * stride;
            if (e < n) __stcs(out + e, (typename Cfg::WO)poly_ct<Cfg, I>((I)a[u]));
        }
    }
}

}  // namespace qb
