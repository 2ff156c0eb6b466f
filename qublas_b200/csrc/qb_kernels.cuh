// qb_kernels.cuh — runtime-descriptor kernels: correct for every plan inside the parity domain.
//
// The programs (qb_prog.h) arrive as __grid_constant__ parameters, i.e. in the constant bank, and
// are uniform across the grid; quantizer modes are uniform branches.  The compile-time-specialised
// fast paths (qb_fast.cuh) instantiate the same primitives with CT = true.
#pragma once
#include "qb_device.cuh"

namespace qb {

// ----------------------------------------------------------------------------------------------
// Qreduce, vector form (QuBLAS.h:4960-4990), data-parallel: one warp reduces a chunk of 2^pw
// consecutive input nodes of one vector into one node `pw` levels up.
//   lane   : 8 contiguous nodes            -> 3 levels in registers
//   warp   : 256 contiguous nodes          -> 5 xor-shuffle levels (lane pairs 1,2,4,8,16 = adjacent pairs)
//   chunk  : 2^(pw-8) such groups in order -> level-indexed stack, replicated in all lanes
// A node is present iff its first leaf index < n_in; a present node without a right sibling is
// carried up through the level's converting assignment (:4977-4980).  Passes are chained with
// lev0 += pw (tree separability) until one node per vector is left.
// ----------------------------------------------------------------------------------------------
template <typename WT, typename I>
QB_DEV void load8(const WT *p, bool vec_ok, I v[8])
{
    if (vec_ok) {
        if constexpr (sizeof(WT) == 4) {
            int4 a = __ldg((const int4 *)p), b = __ldg((const int4 *)p + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
            return;
        } else if constexpr (sizeof(WT) == 8) {
            const longlong2 *q = (const longlong2 *)p;
#pragma unroll
            for (int i = 0; i < 4; i++) { longlong2 t = __ldg(q + i); v[2 * i] = (I)t.x; v[2 * i + 1] = (I)t.y; }
            return;
        } else if constexpr (sizeof(WT) == 2) {
            int4 a = __ldg((const int4 *)p);
            int w[4] = { a.x, a.y, a.z, a.w };
#pragma unroll
            for (int i = 0; i < 4; i++) { v[2 * i] = (I)(int16_t)(w[i] & 0xffff); v[2 * i + 1] = (I)(int16_t)(w[i] >> 16); }
            return;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = (I)__ldg(p + i);
}

template <typename I> QB_DEV I shfl_xor_i(I v, int m)
{
    if constexpr (sizeof(I) == 8) return (I)__shfl_xor_sync(0xffffffffu, (long long)v, m);
    else return (I)__shfl_xor_sync(0xffffffffu, (int)v, m);
}

template <typename I>
QB_DEV void comb(I &lv, bool &lp, I rv, bool rp, uint32_t r, uint32_t rmax, uint32_t lev0, const TreeProg &P)
{
    if (r >= rmax || !lp) return;                       // above the root, or nothing here
    if (rp) lv = merge_apply<I, false>(lv, rv, P.lvl[lev0 + r]);
    else    lv = conv_apply<I, false>(lv, P.carry[lev0 + r]);
}

template <typename WT, typename I>
__global__ void __launch_bounds__(256)
k_reduce_warp(const WT *__restrict__ x, void *__restrict__ out, int out_wb, uint64_t n_in, uint64_t batch,
              uint32_t lev0, uint32_t pw, uint32_t rmax, int vec_ok, const __grid_constant__ TreeProg P)
{
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t per_vec = (n_in + ((uint64_t)1 << pw) - 1) >> pw;
    if (warp >= batch * per_vec) return;
    const uint64_t b = warp / per_vec, c = warp - b * per_vec;
    const WT *xv = x + b * n_in;
    const uint64_t e_begin = c << pw;
    const uint64_t e_end = (e_begin + ((uint64_t)1 << pw)) < n_in ? e_begin + ((uint64_t)1 << pw) : n_in;
    const uint32_t nsub = 1u << (pw - 8);
    I st[32];
    I result;
    if (e_end - e_begin == ((uint64_t)1 << pw) && rmax >= pw) {
        // ---- fast path: every node present, 2^pw leaves ----
        constexpr int U = 4;
        for (uint32_t j0 = 0; j0 < nsub; j0 += U) {
            I v[U][8];
#pragma unroll
            for (int u = 0; u < U; u++)
                if (j0 + u < nsub) load8<WT, I>(xv + e_begin + (uint64_t)(j0 + u) * 256 + lane * 8, vec_ok, v[u]);
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (j0 + u >= nsub) break;
                I *w = v[u];
                I a0 = merge_apply<I, false>(w[0], w[1], P.lvl[lev0]);
                I a1 = merge_apply<I, false>(w[2], w[3], P.lvl[lev0]);
                I a2 = merge_apply<I, false>(w[4], w[5], P.lvl[lev0]);
                I a3 = merge_apply<I, false>(w[6], w[7], P.lvl[lev0]);
                I b0 = merge_apply<I, false>(a0, a1, P.lvl[lev0 + 1]);
                I b1 = merge_apply<I, false>(a2, a3, P.lvl[lev0 + 1]);
                I cur = merge_apply<I, false>(b0, b1, P.lvl[lev0 + 2]);
#pragma unroll
                for (int s = 0; s < 5; s++)
                    cur = merge_apply<I, false>(cur, shfl_xor_i<I>(cur, 1 << s), P.lvl[lev0 + 3 + s]);
                uint32_t l = 0, m = j0 + u;
                while (m & 1) { cur = merge_apply<I, false>(st[l], cur, P.lvl[lev0 + 8 + l]); ++l; m >>= 1; }
                st[l] = cur;
            }
        }
        result = st[pw - 8];
    } else {
        // ---- tail / final chunk: presence-tracked ----
        bool sp[32];
        for (uint32_t j = 0; j < nsub; j++) {
            const uint64_t e0 = e_begin + (uint64_t)j * 256 + lane * 8;
            I v[8]; bool p[8];
#pragma unroll
            for (int i = 0; i < 8; i++) { p[i] = e0 + i < e_end; v[i] = p[i] ? (I)__ldg(xv + e0 + i) : (I)0; }
#pragma unroll
            for (int i = 0; i < 4; i++) comb<I>(v[2 * i], p[2 * i], v[2 * i + 1], p[2 * i + 1], 0, rmax, lev0, P);
            comb<I>(v[0], p[0], v[2], p[2], 1, rmax, lev0, P);
            comb<I>(v[4], p[4], v[6], p[6], 1, rmax, lev0, P);
            comb<I>(v[0], p[0], v[4], p[4], 2, rmax, lev0, P);
            I cur = v[0]; bool cp = p[0];
#pragma unroll
            for (int s = 0; s < 5; s++) {
                I ov = shfl_xor_i<I>(cur, 1 << s);
                bool op = __shfl_xor_sync(0xffffffffu, (int)cp, 1 << s) != 0;
                bool upper = (lane >> s) & 1;
                I lv = upper ? ov : cur, rv = upper ? cur : ov;
                bool lp = upper ? op : cp, rp = upper ? cp : op;
                comb<I>(lv, lp, rv, rp, 3 + s, rmax, lev0, P);
                cur = lv; cp = lp;
            }
            uint32_t l = 0, m = j;
            while (m & 1) {
                I lv = st[l]; bool lp = sp[l];
                comb<I>(lv, lp, cur, cp, 8 + l, rmax, lev0, P);
                cur = lv; cp = lp; ++l; m >>= 1;
            }
            st[l] = cur; sp[l] = cp;
        }
        result = st[pw - 8];
    }
    if (lane == 0) store_word<I>(out, warp, out_wb, result);
}

// ----------------------------------------------------------------------------------------------
// One thread per vector: both tree forms, any length (small vectors / variadic form).
// ----------------------------------------------------------------------------------------------
template <typename I>
__global__ void __launch_bounds__(128)
k_reduce_thread(const void *__restrict__ x, int in_wb, void *__restrict__ out, int out_wb, uint64_t len,
                uint64_t batch, const __grid_constant__ TreeProg P)
{
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    const uint64_t base = b * len;
    I r = tree_stream<I, false, QBP_MAX_LEVELS>(P, len, [&](uint64_t k) { return load_word<I>(x, base + k, in_wb); });
    store_word<I>(out, b, out_wb, r);
}

// Variadic form of a non-power-of-two length, last step (qb_jit.cpp:jit_reduce_var): the passes before it reduced every
// whole 256-node chunk; what is left are runs of nodes at different levels -- the roots of the last pass, then the trailing
// nodes of each pass, highest level first, i.e. in leaf order.  One thread per vector pushes them through the level-indexed
// stack (a node of level L with index idx among the level-L nodes merges while the low bits of idx are set) and runs the
// variadic flush program (QuBLAS.h:4942-4949).
struct VarRun { const int32_t *ptr; uint64_t stride, start, count; uint32_t level, pad; };
struct VarRuns { int n, pad; VarRun r[6]; };
template <typename I>
__global__ void __launch_bounds__(128)
k_reduce_var_finish(const __grid_constant__ VarRuns R, void *__restrict__ out, int out_wb, uint64_t batch, const __grid_constant__ TreeProg P)
{
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    I st[QBP_MAX_LEVELS];
    for (int i = 0; i < R.n; i++) {
        const VarRun &q = R.r[i];
        const int32_t *p = q.ptr + b * q.stride;
        for (uint64_t idx = q.start; idx < q.start + q.count; ++idx) {
            I v = (I)p[idx];
            unsigned l = q.level;
            uint64_t m = idx;
            while (m & 1) { v = merge_apply<I, false>(st[l], v, P.lvl[l]); ++l; m >>= 1; }
            st[l] = v;
        }
    }
    store_word<I>(out, b, out_wb, flush_apply<I, false>(P, st));
}

// ----------------------------------------------------------------------------------------------
// Qgemul, one thread per output element (any shape, transposes, both tree forms).
// C[i,j] = conv_C( tree_k( Qmul<mul>(opA(i,k), opB(k,j)) ) )      readme.md:80-87
// ----------------------------------------------------------------------------------------------
struct GemulShape {
    uint32_t M, N, K;
    uint32_t trans_a, trans_b;
    int32_t  a_wb, b_wb, c_wb;
    int64_t  stride_a, stride_b, stride_c;
};

template <typename I>
__global__ void __launch_bounds__(128)
k_gemul_thread(const void *__restrict__ A, const void *__restrict__ B, void *__restrict__ C, uint64_t batch,
               const __grid_constant__ GemulShape S, const __grid_constant__ QStage mul,
               const __grid_constant__ TreeProg P, const __grid_constant__ QStage outq)
{
    const uint64_t per = (uint64_t)S.M * S.N;
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= batch * per) return;
    const uint64_t inst = t / per, ij = t - inst * per;
    const uint32_t j = (uint32_t)(ij / S.M), i = (uint32_t)(ij - (uint64_t)j * S.M);
    const uint64_t oa = inst * (uint64_t)S.stride_a, ob = inst * (uint64_t)S.stride_b;
    I r = tree_stream<I, false, QBP_MAX_LEVELS>(P, S.K, [&](uint64_t k) {
        const uint64_t ia = S.trans_a ? (k + (uint64_t)i * S.K) : (i + k * S.M);      // column-major, QuBLAS.h:2680-2692
        const uint64_t ib = S.trans_b ? (j + k * S.N) : (k + (uint64_t)j * S.K);
        I a = load_word<I>(A, oa + ia, S.a_wb), b = load_word<I>(B, ob + ib, S.b_wb);
        return mul_apply<I, false>(a, b, mul);
    });
    store_word<I>(C, inst * (uint64_t)S.stride_c + ij, S.c_wb, conv_apply<I, false>(r, outq));
}

// Complex Qgemul with TFComplexMul products: two level-indexed stacks (real / imaginary part),
// complex Qadd is component-wise (QuBLAS.h:3549-3568).
template <typename I>
__global__ void __launch_bounds__(128)
k_cgemul_thread(const void *__restrict__ A, const void *__restrict__ B, void *__restrict__ C, uint64_t batch,
                const __grid_constant__ GemulShape S, const int kind, const __grid_constant__ TfProg tf, const __grid_constant__ BcmProg bcm,
                const __grid_constant__ TreeProg Pre, const __grid_constant__ TreeProg Pim,
                const __grid_constant__ QStage out_re, const __grid_constant__ QStage out_im)
{
    const uint64_t per = (uint64_t)S.M * S.N;
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= batch * per) return;
    const uint64_t inst = t / per, ij = t - inst * per;
    const uint32_t j = (uint32_t)(ij / S.M), i = (uint32_t)(ij - (uint64_t)j * S.M);
    const uint64_t oa = inst * (uint64_t)S.stride_a, ob = inst * (uint64_t)S.stride_b;
    I sr[QBP_MAX_LEVELS], si[QBP_MAX_LEVELS];
    for (uint64_t k = 0; k < S.K; ++k) {
        const uint64_t ia = oa + (S.trans_a ? (k + (uint64_t)i * S.K) : (i + k * S.M));
        const uint64_t ib = ob + (S.trans_b ? (j + k * S.N) : (k + (uint64_t)j * S.K));
        I a, b = 0, c, d = 0;
        if (kind == 2) a = load_word<I>(A, ia, S.a_wb);                        // real A: one word per element
        else { a = load_word<I>(A, 2 * ia, S.a_wb); b = load_word<I>(A, 2 * ia + 1, S.a_wb); }
        if (kind == 3) c = load_word<I>(B, ib, S.b_wb);                        // real B
        else { c = load_word<I>(B, 2 * ib, S.b_wb); d = load_word<I>(B, 2 * ib + 1, S.b_wb); }
        I vr, vi;
        cprod_apply<I, false>(kind, a, b, c, d, tf, bcm, vr, vi);
        unsigned l = 0; uint64_t m = k;
        while (m & 1) {
            vr = merge_apply<I, false>(sr[l], vr, Pre.lvl[l]);
            vi = merge_apply<I, false>(si[l], vi, Pim.lvl[l]);
            ++l; m >>= 1;
        }
        sr[l] = vr; si[l] = vi;
    }
    I rr = flush_apply<I, false>(Pre, sr), ri = flush_apply<I, false>(Pim, si);
    const uint64_t ic = inst * (uint64_t)S.stride_c + ij;
    store_word<I>(C, 2 * ic, S.c_wb, conv_apply<I, false>(rr, out_re));
    store_word<I>(C, 2 * ic + 1, S.c_wb, conv_apply<I, false>(ri, out_im));
}

// ----------------------------------------------------------------------------------------------
// Element-wise kernels
// ----------------------------------------------------------------------------------------------
template <typename I>
__global__ void k_tfcmul(const void *__restrict__ x, const void *__restrict__ y, void *__restrict__ out, uint64_t n,
                         int x_wb, int y_wb, int o_wb, const int kind, const __grid_constant__ TfProg tf, const __grid_constant__ BcmProg bcm)
{
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (uint64_t)gridDim.x * blockDim.x) {
        I a, b = 0, c, d = 0;
        if (kind == 2) a = load_word<I>(x, e, x_wb);
        else { a = load_word<I>(x, 2 * e, x_wb); b = load_word<I>(x, 2 * e + 1, x_wb); }
        if (kind == 3) c = load_word<I>(y, e, y_wb);
        else { c = load_word<I>(y, 2 * e, y_wb); d = load_word<I>(y, 2 * e + 1, y_wb); }
        I re, im;
        cprod_apply<I, false>(kind, a, b, c, d, tf, bcm, re, im);
        store_word<I>(out, 2 * e, o_wb, re);
        store_word<I>(out, 2 * e + 1, o_wb, im);
    }
}

template <typename I>
__global__ void k_convert(const void *__restrict__ in, void *__restrict__ out, uint64_t n, int in_wb, int out_wb,
                          const __grid_constant__ QStage q)
{
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (uint64_t)gridDim.x * blockDim.x)
        store_word<I>(out, e, out_wb, conv_apply<I, false>(load_word<I>(in, e, in_wb), q));
}

// element-wise Qmul / Qadd / Qsub (op 0 / 1 / 2)
template <typename I>
__global__ void k_elementwise(const void *__restrict__ a, const void *__restrict__ b, void *__restrict__ out, uint64_t n, int a_wb, int b_wb,
                              int o_wb, int op, const __grid_constant__ AddStage st)
{
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (uint64_t)gridDim.x * blockDim.x) {
        const I x = load_word<I>(a, e, a_wb), y = load_word<I>(b, e, b_wb);
        const I r = op == 0 ? mul_apply<I, false>(x, y, st.q) : add_apply<I, false>(x, y, st);
        store_word<I>(out, e, o_wb, r);
    }
}

// ANUS::Qpoly / Qapprox over a tensor; in place is fine (one element per thread and step)
template <typename I>
__global__ void __launch_bounds__(256) k_poly(const void *__restrict__ x, void *__restrict__ out, uint64_t n, int x_wb, int o_wb, const __grid_constant__ PolyProg P)
{
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (uint64_t)gridDim.x * blockDim.x)
        store_word<I>(out, e, o_wb, poly_apply<I, false>(P, load_word<I>(x, e, x_wb)));
}

// ANUS::Qtable: out[e] = table[raw(x[e]) - raw_min]
__global__ void __launch_bounds__(256) k_table(const void *__restrict__ x, void *__restrict__ out, const void *__restrict__ table, uint64_t n, int x_wb, int o_wb,
                                               int64_t raw_min, uint64_t entries)
{
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t idx = (uint64_t)(load_word<int64_t>(x, e, x_wb) - raw_min);
        if (idx >= entries) idx = entries - 1;                 // words outside the type's range cannot occur in a well-formed buffer
        store_word<int64_t>(out, e, o_wb, load_word<int64_t>(table, idx, o_wb));
    }
}

__global__ void k_synth_fill(void *__restrict__ dst, uint64_t count, int wb, int64_t lo, uint64_t span, uint64_t key,
                             uint64_t first_index)
{
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (uint64_t)gridDim.x * blockDim.x)
        store_word<int64_t>(dst, e, wb, synth_word(key, first_index + e, lo, span));
}

// Widen / narrow raw words (used to hand int64 partials to the caller).
__global__ void k_recast(const void *__restrict__ in, int in_wb, void *__restrict__ out, int out_wb, uint64_t n)
{
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (uint64_t)gridDim.x * blockDim.x)
        store_word<int64_t>(out, e, out_wb, load_word<int64_t>(in, e, in_wb));
}

}  // namespace qb
