// qb_fast_gemul.cuh — compile-time-specialised batched Qgemul for small dense instances
// (BASELINE config 3: 10^6 x 64x64x64, 16-bit operands) on the integer ALU.
//
//   * one CTA = one instance at a time, persistent over the batch (grid = resident CTAs x SMs);
//   * A and B instances (contiguous column-major arrays) are staged in shared memory by 1-D TMA
//     bulk copies (cp.async.bulk ... mbarrier::complete_tx) into a 2-stage ring, so the next
//     instance streams in from HBM while the current one is computed;
//   * each thread owns a TM x TN register tile of C and consumes k in order: 8 products are
//     reduced by a fully unrolled 3-level adjacent-pair tree, the resulting level-3 nodes go
//     through a level-indexed register stack (uniform branches on the chunk counter), which is
//     exactly the reference's pairwise order (QuBLAS.h:4936-4941) with its per-level quantizer;
//   * every shift / bound / mode of Qmul<MulArgs> (QuBLAS.h:3146-3171), of the per-level
//     Qadd (QuBLAS.h:3177-3204) and of the final converting assignment (QuBLAS.h:2398-2411) is an
//     immediate; clamps the host's interval analysis proved dead are not emitted.
#pragma once
#include "qb_device.cuh"

namespace qb {

QB_DEV uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

QB_DEV void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
QB_DEV void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
QB_DEV void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
QB_DEV void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <class Cfg, int TM, int TN>
struct GemulTile {
    typedef typename Cfg::I I;
    static constexpr int M = Cfg::M, N = Cfg::N, K = Cfg::K;
    static constexpr int KC = 8;                               // products reduced in registers per chunk
    static constexpr int NCH = K / KC;
    static constexpr int NTOP = Cfg::NLEV - 3;                 // stack levels above a chunk
    static constexpr int G = Cfg::G;                             // instances a CTA works on at a time (small shapes: fill the block)
    static constexpr int TPI = (M / TM) * (N / TN);              // threads per instance, each owns a 4x4 tile
    static constexpr int THREADS = G * TPI;                      // threads that own a tile
    static constexpr int BLOCK = (THREADS + 31) / 32 * 32;       // launched threads: the rest only helps with staging, conversion and barriers
    static constexpr int A_WORDS = M * K, B_WORDS = K * N;       // per instance
    // operand words: int32 (the reference's storage) or packed int16 (qb_qdesc.word_bytes = 2: half the HBM / PCIe bytes).
    // Packed words are staged as they are and widened by the conversion pass into a separate single-buffered region.
    static constexpr int WORD = (int)sizeof(typename Cfg::WA);
    static constexpr bool NARROW = WORD == 2;
    static constexpr uint32_t STAGE_BYTES = G * (A_WORDS + B_WORDS) * WORD;
    static constexpr uint32_t WIDE_BYTES = NARROW ? G * (A_WORDS + B_WORDS) * 4 : 0;
    static constexpr uint32_t SMEM_BYTES = 2 * STAGE_BYTES + WIDE_BYTES + 64;
    static_assert(M % TM == 0 && N % TN == 0 && TM == 4 && TN == 4, "tile shape");
    static constexpr bool K_POW2 = (K & (K - 1)) == 0;         // otherwise the pending nodes go through the flush program
    static_assert(K % KC == 0 && K >= KC, "K must be a multiple of 8");
    static_assert(sizeof(typename Cfg::WA) == sizeof(typename Cfg::WB) && (WORD == 4 || WORD == 2) && sizeof(typename Cfg::WC) == 4, "int32 or int16 operand words, int32 results");
    static_assert(!NARROW || (A_WORDS % 8 == 0 && B_WORDS % 8 == 0), "16-byte aligned int16 instances");
    // batch strides in words (plan.stride_*): dense by default, 0 = the operand is shared by all instances, larger = padded
    static constexpr unsigned long long SA = Cfg::SA, SB = Cfg::SB, SC = Cfg::SC;
    static_assert(G == 1 || (SA == (unsigned long long)A_WORDS && SB == (unsigned long long)B_WORDS), "groups of instances need dense operands");
    static_assert(SA % (16 / WORD) == 0 && SB % (16 / WORD) == 0 && SC % 4 == 0, "16-byte aligned instances");
};

// three-input min / max (one VIMNMX3 each on sm_90+)
QB_DEV int min3i(int a, int b, int c) { return __vimin3_s32(a, b, c); }
QB_DEV int max3i(int a, int b, int c) { return __vimax3_s32(a, b, c); }

// level-indexed register stack above the 8-product chunks; c = chunk counter (uniform)
template <class Cfg, int L, int NTOP, typename I>
QB_DEV void stack_push(I (&st)[NTOP > 0 ? NTOP : 1], uint32_t c, I v, I &result)
{
    if constexpr (L >= NTOP) {
        result = v;
    } else {
        if ((c >> L) & 1) {
            constexpr AddStage lv = Cfg::lvl(3 + L);
            I m = merge_apply<I, true>(st[L], v, lv);
            stack_push<Cfg, L + 1, NTOP, I>(st, c, m, result);
        } else {
            st[L] = v;
        }
    }
}

// Occupancy target: 512 resident threads per SM at the 128-register cap, bounded by shared memory.  Measured on B200
// (config 3, via QB200_JIT_DEFINES=-DQB_MIN_BLOCKS=n): 2 CTAs of 256 threads (16 warps, 52 words spilled) and 1 CTA with
// 202 registers and no spills run within noise of each other, 3 CTAs (85 registers) are 6 % slower, and prefetching the
// operands one half-chunk ahead (235 registers) changes nothing -- the loop is bound by pipe issue, not by latency.
constexpr int gemul_min_blocks(int threads, unsigned smem_bytes)
{
#ifdef QB_MIN_BLOCKS
    (void)threads; (void)smem_bytes;
    return QB_MIN_BLOCKS;
#else
    int n = 512 / threads, by_smem = (int)((227u * 1024u) / (smem_bytes + 1024u));
    if (n > by_smem) n = by_smem;
    return n < 1 ? 1 : (n > 8 ? 8 : n);
#endif
}

// K = 8 * (not a power of two): the same stack, but nodes stay pending at the end and are combined by the host-built
// flush program (vector form QuBLAS.h:4977-4980, variadic form :4942-4949), unrolled at compile time.  Slot = level - 3.
template <class Cfg, int L, int NTOP, typename I>
QB_DEV void stack_push_pending(I (&st)[NTOP > 0 ? NTOP : 1], uint32_t c, I v)
{
    if constexpr (L < NTOP) {
        if ((c >> L) & 1) {
            constexpr AddStage lv = Cfg::lvl(3 + L);
            stack_push_pending<Cfg, L + 1, NTOP, I>(st, c, merge_apply<I, true>(st[L], v, lv));
        } else {
            st[L] = v;
        }
    }
}
template <class Cfg, int IDX, int NTOP, typename I>
QB_DEV I stack_flush(const I (&st)[NTOP > 0 ? NTOP : 1], I cur)
{
    if constexpr (IDX >= Cfg::NFLUSH) {
        return cur;
    } else {
        constexpr FlushStep f = Cfg::flush(IDX);
        if constexpr (f.kind == FL_LOAD) cur = st[f.slot - 3];
        else if constexpr (f.kind == FL_CONV) cur = conv_apply<I, true>(cur, f.a.q);
        else cur = add_apply<I, true>(st[f.slot - 3], cur, f.a);
        return stack_flush<Cfg, IDX + 1, NTOP, I>(st, cur);
    }
}

constexpr QStage stage_without_clamps(QStage q) { q.need_lo = 0; q.need_hi = 0; return q; }

// One instance's 4x4 register tile.  CLAMP = false is the same arithmetic with Qmul's saturation removed;
// it is selected per instance when the instance's own operand ranges prove that no product can saturate.
template <class Cfg, bool CLAMP>
QB_DEV void tile_compute(const int32_t *As, const int32_t *Bs, int i0, int j0, typename Cfg::I (&res)[4][4])
{
    typedef GemulTile<Cfg, 4, 4> T;
    typedef typename T::I I;
    constexpr int M = T::M, K = T::K, NTOP = T::NTOP;
    constexpr QStage mulq = CLAMP ? Cfg::mul() : stage_without_clamps(Cfg::mul());
    constexpr AddStage l0 = Cfg::lvl(0), l1 = Cfg::lvl(1), l2 = Cfg::lvl(2);
    const int4 *As4 = (const int4 *)As;
    const int4 *Bs4 = (const int4 *)Bs;
    I st[4][4][NTOP > 0 ? NTOP : 1];
    // The chunk loop is unrolled (up to 8 chunks = the whole K = 64 of config 3): the branches of the level-indexed stack
    // on the chunk counter then fold at compile time, the stack slots get exact live ranges instead of being loop-carried,
    // and the scheduler can overlap the tail of one chunk with the loads of the next.  Measured on config 3 (NVRTC build):
    // unroll 1 / 2 / 4 / 8 = 8 398 / 8 978 / 8 929 / 10 244 GMAC/s.  QB_CHUNK_UNROLL overrides (kernel experiments).
#ifndef QB_CHUNK_UNROLL
#define QB_CHUNK_UNROLL 0
#endif
    // (not for blocks above 512 threads: their register cap is 64 and the unrolled body spills: 128x128x64 drops from 8.3 to 6.7 TMAC/s)
    constexpr int CHUNK_UNROLL = QB_CHUNK_UNROLL > 0 ? QB_CHUNK_UNROLL : (T::BLOCK > 512 ? 1 : (T::NCH <= 8 ? T::NCH : 8));
#pragma unroll CHUNK_UNROLL
    for (uint32_t c = 0; c < (uint32_t)T::NCH; ++c) {
        I half0[4][4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k0 = c * 8 + h * 4;
            // a 4 (i) x 4 (k) block of op(A) and a 4 (k) x 4 (j) block of op(B) as four 16-byte loads each; which index runs
            // inside a load depends on the storage: A is M x K column-major (i contiguous), transposed K x M (k contiguous);
            // B is K x N column-major (k contiguous), transposed N x K (j contiguous)
            int4 bq[4], aq[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) bq[u] = Cfg::TRANS_B ? Bs4[((k0 + u) * T::N + j0) >> 2] : Bs4[((j0 + u) * K + k0) >> 2];
#pragma unroll
            for (int u = 0; u < 4; ++u) aq[u] = Cfg::TRANS_A ? As4[((i0 + u) * K + k0) >> 2] : As4[((k0 + u) * M + i0) >> 2];
            auto a_at = [&](int ii, int kk) -> I { return Cfg::TRANS_A ? (&aq[ii].x)[kk] : (&aq[kk].x)[ii]; };
            auto b_at = [&](int kk, int jj) -> I { return Cfg::TRANS_B ? (&bq[kk].x)[jj] : (&bq[jj].x)[kk]; };
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const I b0 = b_at(0, jj), b1 = b_at(1, jj), b2 = b_at(2, jj), b3 = b_at(3, jj);
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) {
                    const I a0 = a_at(ii, 0), a1 = a_at(ii, 1), a2 = a_at(ii, 2), a3 = a_at(ii, 3);
                    I s01, s23;
                    if constexpr (Cfg::FP_MUL) {
                        constexpr int S = Cfg::FP_SHIFT;
                        constexpr uint32_t KB2 = 2u * (uint32_t)fp_kbits<S>();
                        const int p0 = fp_mul_apply<S, true>(__int_as_float(a0), __int_as_float(b0), mulq);
                        const int p1 = fp_mul_apply<S, true>(__int_as_float(a1), __int_as_float(b1), mulq);
                        const int p2 = fp_mul_apply<S, true>(__int_as_float(a2), __int_as_float(b2), mulq);
                        const int p3 = fp_mul_apply<S, true>(__int_as_float(a3), __int_as_float(b3), mulq);
                        s01 = q_apply<I, true>((I)((uint32_t)p0 + (uint32_t)p1 - KB2), l0.q);       // unbias folded into the level-0 add
                        s23 = q_apply<I, true>((I)((uint32_t)p2 + (uint32_t)p3 - KB2), l0.q);
                    } else {
                        I p0 = mul_apply<I, true>(a0, b0, mulq);
                        I p1 = mul_apply<I, true>(a1, b1, mulq);
                        I p2 = mul_apply<I, true>(a2, b2, mulq);
                        I p3 = mul_apply<I, true>(a3, b3, mulq);
                        s01 = merge_apply<I, true>(p0, p1, l0);
                        s23 = merge_apply<I, true>(p2, p3, l0);
                    }
                    const I hv = merge_apply<I, true>(s01, s23, l1);
                    if (h == 0) {
                        half0[ii][jj] = hv;
                    } else {
                        const I v = merge_apply<I, true>(half0[ii][jj], hv, l2);
                        if constexpr (T::K_POW2) stack_push<Cfg, 0, NTOP, I>(st[ii][jj], c, v, res[ii][jj]);
                        else stack_push_pending<Cfg, 0, NTOP, I>(st[ii][jj], c, v);
                    }
                }
            }
        }
    }
    if constexpr (!T::K_POW2) {
#pragma unroll
        for (int ii = 0; ii < 4; ++ii)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) res[ii][jj] = stack_flush<Cfg, 0, NTOP, I>(st[ii][jj], (I)0);
    }
}

template <class Cfg>
QB_DEV void gemul_tile_body(const typename Cfg::WA *__restrict__ A, const typename Cfg::WB *__restrict__ B, int32_t *__restrict__ C, uint64_t batch)
{
    typedef GemulTile<Cfg, 4, 4> T;
    typedef typename T::I I;
    constexpr int M = T::M;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bars = (uint64_t *)(smem_raw + 2 * T::STAGE_BYTES + T::WIDE_BYTES);
    __shared__ int range_smem[32][4];

    const int tid = threadIdx.x;
    const int g = T::G == 1 ? 0 : tid / T::TPI, tl = T::G == 1 ? tid : tid - g * T::TPI;       // instance within the group, thread within the instance
    const int tx = tl % (M / 4), ty = tl / (M / 4);
    const int i0 = tx * 4, j0 = ty * 4;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // a work item is a group of G consecutive instances (the last group of the batch may be shorter); their A arrays are
    // contiguous in memory, and so are their B arrays: two bulk copies per group
    const uint64_t ngroups = (batch + T::G - 1) / T::G;
    auto group_count = [&](uint64_t grp) -> uint32_t {
        const uint64_t left = batch - grp * T::G;
        return (uint32_t)(left < (uint64_t)T::G ? left : (uint64_t)T::G);
    };
    auto issue = [&](uint64_t grp, int s) {
        unsigned char *dst = smem_raw + (size_t)s * T::STAGE_BYTES;
        const uint32_t cnt = group_count(grp);
        mbar_expect_tx(&bars[s], cnt * (uint32_t)(T::A_WORDS + T::B_WORDS) * (uint32_t)T::WORD);
        tma_load_1d(dst, A + grp * T::G * T::SA, cnt * (uint32_t)T::A_WORDS * (uint32_t)T::WORD, &bars[s]);
        tma_load_1d(dst + T::G * T::A_WORDS * T::WORD, B + grp * T::G * T::SB, cnt * (uint32_t)T::B_WORDS * (uint32_t)T::WORD, &bars[s]);
    };

    uint64_t inst = blockIdx.x;                                   // group index
    if (inst < ngroups && tid == 0) issue(inst, 0);

    constexpr QStage mulq = Cfg::mul();
    constexpr QStage outq = Cfg::out();
    // Dynamic clamp elision: Qmul's SAT clamp survives the static interval analysis only because of extreme
    // operand values (config 3: min x min).  One pass over the staged instance finds its actual operand
    // ranges; if even the extreme corner products stay inside [P_LO, P_HI] (host-computed, conservative),
    // the clamp-free instantiation runs.  Same bits either way.
    constexpr bool DYN = (mulq.need_lo || mulq.need_hi) && (mulq.of == SAT_TCPL || mulq.of == SAT_SMGN);
    constexpr bool PASS = DYN || Cfg::FP_MUL || T::NARROW;

    for (uint32_t it = 0; inst < ngroups; inst += gridDim.x, ++it) {
        const int s = it & 1;
        const uint64_t next = inst + gridDim.x;
        if (tid == 0 && next < ngroups) {
            if constexpr (Cfg::FP_MUL) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes (int->float) before TMA refill
            issue(next, s ^ 1);                                   // stage s^1 was released by the barrier below
        }
        mbar_wait(&bars[s], (it >> 1) & 1);
        unsigned char *stage = smem_raw + (size_t)s * T::STAGE_BYTES;
        int32_t *As = T::NARROW ? (int32_t *)(smem_raw + 2 * T::STAGE_BYTES) : (int32_t *)stage;      // 32-bit operands the tiles read
        int32_t *Bs = As + T::G * T::A_WORDS;
        bool clamp = true;
        if constexpr (PASS) {
            // one pass over the staged operands: exact int->float conversion (FP_MUL) and per-instance ranges (DYN)
            int lo[2] = { 0x7fffffff, 0x7fffffff }, hi[2] = { (int)0x80000000, (int)0x80000000 };
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                int4 *w = (int4 *)(part ? Bs : As);
                if constexpr (T::NARROW) {
                    // packed int16 words: widen (sign-extend) into the 32-bit region, converting / scanning on the way
                    const int4 *raw = (const int4 *)(stage + (part ? T::G * T::A_WORDS * 2 : 0));
                    const int n8 = (part ? T::G * T::B_WORDS : T::G * T::A_WORDS) / 8;
                    for (int i = tid; i < n8; i += T::BLOCK) {
                        const int4 v = raw[i];
                        const int e[8] = { (int)(short)(v.x & 0xffff), v.x >> 16, (int)(short)(v.y & 0xffff), v.y >> 16,
                                           (int)(short)(v.z & 0xffff), v.z >> 16, (int)(short)(v.w & 0xffff), v.w >> 16 };
                        if constexpr (DYN) {
                            lo[part] = min3i(min3i(min3i(min3i(lo[part], e[0], e[1]), e[2], e[3]), e[4], e[5]), e[6], e[7]);
                            hi[part] = max3i(max3i(max3i(max3i(hi[part], e[0], e[1]), e[2], e[3]), e[4], e[5]), e[6], e[7]);
                        }
                        if constexpr (Cfg::FP_MUL) {
                            ((float4 *)w)[2 * i] = make_float4(__int2float_rn(e[0]), __int2float_rn(e[1]), __int2float_rn(e[2]), __int2float_rn(e[3]));
                            ((float4 *)w)[2 * i + 1] = make_float4(__int2float_rn(e[4]), __int2float_rn(e[5]), __int2float_rn(e[6]), __int2float_rn(e[7]));
                        } else {
                            w[2 * i] = make_int4(e[0], e[1], e[2], e[3]);
                            w[2 * i + 1] = make_int4(e[4], e[5], e[6], e[7]);
                        }
                    }
                    continue;
                }
                constexpr int N4A = T::G * T::A_WORDS / 4, N4B = T::G * T::B_WORDS / 4;      // the whole group (a short last group leaves stale words: harmless, conservative for the ranges)
                const int n4 = part ? N4B : N4A;
                constexpr int PER = ((N4A > N4B ? N4A : N4B) + T::BLOCK - 1) / T::BLOCK;
                if constexpr (PER <= 8) {
                    // all loads of this part in flight first, then ranges / conversion / stores
                    int4 v[PER];
#pragma unroll
                    for (int u = 0; u < PER; ++u) { const int i = tid + u * T::BLOCK; if (i < n4) v[u] = w[i]; }
#pragma unroll
                    for (int u = 0; u < PER; ++u) {
                        const int i = tid + u * T::BLOCK;
                        if (i < n4) {
                            if constexpr (DYN) {
                                lo[part] = min3i(min3i(lo[part], v[u].x, v[u].y), v[u].z, v[u].w);      // two VIMNMX3 per four words
                                hi[part] = max3i(max3i(hi[part], v[u].x, v[u].y), v[u].z, v[u].w);
                            }
                            if constexpr (Cfg::FP_MUL)
                                ((float4 *)w)[i] = make_float4(__int2float_rn(v[u].x), __int2float_rn(v[u].y), __int2float_rn(v[u].z), __int2float_rn(v[u].w));
                        }
                    }
                } else {
                    for (int i = tid; i < n4; i += T::BLOCK) {
                        const int4 v = w[i];
                        if constexpr (DYN) {
                            lo[part] = min3i(min3i(lo[part], v.x, v.y), v.z, v.w);
                            hi[part] = max3i(max3i(hi[part], v.x, v.y), v.z, v.w);
                        }
                        if constexpr (Cfg::FP_MUL)
                            ((float4 *)w)[i] = make_float4(__int2float_rn(v.x), __int2float_rn(v.y), __int2float_rn(v.z), __int2float_rn(v.w));
                    }
                }
            }
            if constexpr (DYN) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    lo[0] = imin<int>(lo[0], __shfl_xor_sync(0xffffffffu, lo[0], o)); hi[0] = imax<int>(hi[0], __shfl_xor_sync(0xffffffffu, hi[0], o));
                    lo[1] = imin<int>(lo[1], __shfl_xor_sync(0xffffffffu, lo[1], o)); hi[1] = imax<int>(hi[1], __shfl_xor_sync(0xffffffffu, hi[1], o));
                }
                if ((tid & 31) == 0) { range_smem[tid >> 5][0] = lo[0]; range_smem[tid >> 5][1] = hi[0]; range_smem[tid >> 5][2] = lo[1]; range_smem[tid >> 5][3] = hi[1]; }
            }
            __syncthreads();
            if constexpr (DYN) {
                int alo = 0x7fffffff, ahi = (int)0x80000000, blo = 0x7fffffff, bhi = (int)0x80000000;
#pragma unroll
                for (int wv = 0; wv < T::BLOCK / 32; ++wv) {
                    alo = imin<int>(alo, range_smem[wv][0]); ahi = imax<int>(ahi, range_smem[wv][1]);
                    blo = imin<int>(blo, range_smem[wv][2]); bhi = imax<int>(bhi, range_smem[wv][3]);
                }
                const long long c0 = (long long)alo * blo, c1 = (long long)alo * bhi, c2 = (long long)ahi * blo, c3 = (long long)ahi * bhi;
                const long long pmax = imax<long long>(imax<long long>(c0, c1), imax<long long>(c2, c3));
                const long long pmin = imin<long long>(imin<long long>(c0, c1), imin<long long>(c2, c3));
                clamp = (mulq.need_hi && pmax > Cfg::P_HI) || (mulq.need_lo && pmin < Cfg::P_LO);
            }
        }

        if ((T::BLOCK == T::THREADS || tid < T::THREADS) && (T::G == 1 || (uint32_t)g < group_count(inst))) {   // padding threads / a short last group own no tile
            const int32_t *Ag = As + g * T::A_WORDS, *Bg = Bs + g * T::B_WORDS;
            I res[4][4];
            if constexpr (DYN) {
                if (clamp) tile_compute<Cfg, true>(Ag, Bg, i0, j0, res);
                else       tile_compute<Cfg, false>(Ag, Bg, i0, j0, res);
            } else {
                tile_compute<Cfg, true>(Ag, Bg, i0, j0, res);
            }
            // converting assignment into C's element type, column-major 128-bit stores
            int32_t *Ci = C + (inst * T::G + g) * T::SC;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                int4 o;
                o.x = (int32_t)conv_apply<I, true>(res[0][jj], outq);
                o.y = (int32_t)conv_apply<I, true>(res[1][jj], outq);
                o.z = (int32_t)conv_apply<I, true>(res[2][jj], outq);
                o.w = (int32_t)conv_apply<I, true>(res[3][jj], outq);
                *(int4 *)(Ci + (j0 + jj) * M + i0) = o;
            }
        }
        __syncthreads();      // everyone is done reading stage s (and range_smem) before it is refilled
    }
}

template <class Cfg>
__global__ void __launch_bounds__(GemulTile<Cfg, 4, 4>::BLOCK, gemul_min_blocks(GemulTile<Cfg, 4, 4>::BLOCK, GemulTile<Cfg, 4, 4>::SMEM_BYTES))
k_gemul_tile(const typename Cfg::WA *__restrict__ A, const typename Cfg::WB *__restrict__ B, int32_t *__restrict__ C, uint64_t batch)
{
    gemul_tile_body<Cfg>(A, B, C, batch);
}

#ifndef QB_JIT
template <class Cfg>
static int launch_gemul_tile(const void *A, const void *B, void *C, uint64_t batch, cudaStream_t s)
{
    typedef GemulTile<Cfg, 4, 4> T;
    static int grid_caps[64] = { 0 };                 // per device: function attributes and occupancy belong to a context
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    int &grid_cap = grid_caps[cur_dev & 63];
    if (!grid_cap) {
        cudaError_t e = cudaFuncSetAttribute(k_gemul_tile<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gemul_tile<Cfg>, T::BLOCK, T::SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        grid_cap = sms * (per_sm > 0 ? per_sm : 1);
    }
    const uint64_t ngroups = (batch + T::G - 1) / T::G;
    const unsigned grid = (unsigned)(ngroups < (uint64_t)grid_cap ? ngroups : (uint64_t)grid_cap);
    k_gemul_tile<Cfg><<<grid, T::BLOCK, T::SMEM_BYTES, s>>>((const typename Cfg::WA *)A, (const typename Cfg::WB *)B, (int32_t *)C, batch);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) count_launch(1);
    return (int)e;
}

#endif  // QB_JIT

}  // namespace qb
