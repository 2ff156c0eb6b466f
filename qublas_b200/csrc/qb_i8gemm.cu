// qb_i8gemm.cu — tcgen05 kind::i8 tensor-core path of Qgemul (BASELINE config 5).
//
// Used ONLY when the host's interval analysis proves the whole accumulation lossless
// (qb_plan.cpp:gemul_tensor_core_ok): operand raws fit int8, every Qmul<MulArgs> and every tree
// level is an identity, so the reference's tree result IS the exact integer dot product (sum order
// is irrelevant) and fits int32.  The tensor core computes that dot product exactly; the only
// quantizer left is the converting assignment into C's element type (QuBLAS.h:2398-2411), applied
// once in the epilogue.
//
// Pipeline (one CTA per SM, persistent over 128x256 output tiles, warp-specialised):
//   pack   : ABI words (int32/int64, column-major) -> int8, both operands K-major (separate kernels)
//   warp 0 : TMA producer, 2-D tiled loads with 128B swizzle into a 4-stage smem ring (UTMALDG)
//   warp 1 : allocates TMEM (512 columns = two 128x256 int32 accumulators), issues
//            tcgen05.mma.cta_group::1.kind::i8 (M=128, N=256, K=32), commits to mbarriers (UTCQMMA/UTCIMMA)
//   warps 2-5: epilogue, tcgen05.ld 32x32b.x32 (LDTM) -> C's quantizer -> coalesced column-major stores,
//            overlapped with the next tile's MMAs through the second accumulator.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include "qb_device.cuh"
#include "qb_fast.h"
#include "qb_i8gemm.h"

namespace qb {

static constexpr int BM = 128, BN = 256, BK = 128;            // BK in int8 elements = one 128-byte swizzle row
static constexpr int STAGES = 4;
static constexpr int A_STAGE = BM * BK, B_STAGE = BN * BK;    // bytes
static constexpr int STAGE_BYTES = A_STAGE + B_STAGE;
static constexpr int TMEM_COLS = 512;
static constexpr int I8_THREADS = 192;
static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bar_init(uint64_t *b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n)); }
__device__ __forceinline__ void bar_expect_tx(uint64_t *b, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void bar_wait(uint64_t *b, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(s32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(s32(dst)), "l"(map), "r"(s32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, uint16_t mask)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
                 ::"r"(s32(dst)), "l"(map), "r"(s32(bar)), "h"(mask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t n_clusters_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint64_t *bar, uint16_t mask)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(s32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], int8 x int8 -> int32
__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart
// (cute::UMMA::SmemDescriptor: start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version=1 <<46 | SWIZZLE_128B=2 <<61).
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                       // leading byte offset (unused for swizzled K-major), 16-byte units
    d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
    return d;
}
// cute::UMMA::InstrDescriptor for kind::i8: S32 accumulator, signed int8 A and B, both K-major.
__host__ __device__ constexpr uint32_t umma_idesc_i8(int m, int n)
{
    return (2u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack4(int a, int b, int c, int d)
{
    return (uint32_t)(a & 0xff) | ((uint32_t)(b & 0xff) << 8) | ((uint32_t)(c & 0xff) << 16) | ((uint32_t)d << 24);
}

struct I8Params {
    uint32_t M, N, K;             // M, N: real extents of C (leading dimension M, stores are guarded); K: padded to a multiple of BK
    uint32_t tiles_m, tiles_n;
    int32_t  c_wb;
    uint32_t group_m;             // rasterisation: group_m m-tiles share each B panel while it is hot in L2
    uint32_t n_out, row0;         // n_out > 0 (pair kernels): fan the result out into outs[0 .. n_out), rows starting at row0, leading dimension ldc
    uint64_t ldc;
    void    *outs[8];
};

// MC = false: one CTA per tile, plain TMA.
// MC = true : clusters of two CTAs that work on vertically adjacent tiles (same n-tile): each CTA loads its own
//             A tile and ONE HALF of the shared B tile, multicast into both CTAs' shared memory, which cuts the
//             L2->SM traffic per MMA by a third (48 KB -> 32 KB per 128x256x128 block).
template <bool MC>
__device__ __forceinline__ void i8gemm_body(const CUtensorMap &mapA, const CUtensorMap &mapB, void *__restrict__ C, const I8Params &P, const QStage &outq)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);     // SWIZZLE_128B atoms need 1024-byte alignment
    unsigned char *sA = smem, *sB = smem + STAGES * A_STAGE;
    uint64_t *bars = (uint64_t *)(smem + STAGES * STAGE_BYTES);
    uint64_t *full = bars, *empty = bars + STAGES, *tfull = bars + 2 * STAGES, *tempty = bars + 2 * STAGES + 2;
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t nkb = P.K / BK;

    const uint32_t rank = MC ? cluster_ctarank() : 0;
    const uint32_t worker = MC ? cluster_id_x() : blockIdx.x, n_workers = MC ? n_clusters_x() : gridDim.x;
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; s++) { bar_init(&full[s], 1); bar_init(&empty[s], MC ? 2 : 1); }
        for (int a = 0; a < 2; a++) { bar_init(&tfull[a], 1); bar_init(&tempty[a], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (MC) cluster_sync_all();          // peer barriers are initialised before any multicast / remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // work items: tiles (MC = false) or vertical tile pairs (MC = true); rows of work items are grouped so that a
    // B panel stays hot in L2 while `group_m` rows use it
    const uint32_t rows = MC ? P.tiles_m / 2 : P.tiles_m;
    const uint32_t nwork = rows * P.tiles_n;
    auto tile_coords = [&](uint32_t t, uint32_t &tm, uint32_t &tn) {
        const uint32_t per_group = P.group_m * P.tiles_n;
        const uint32_t g = t / per_group, r = t - g * per_group;
        const uint32_t gm = (rows - g * P.group_m) < P.group_m ? (rows - g * P.group_m) : P.group_m;
        const uint32_t row = g * P.group_m + r % gm;
        tm = MC ? 2 * row + rank : row;
        tn = r / gm;
    };

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t t = worker; t < nwork; t += n_workers) {
                uint32_t tm, tn; tile_coords(t, tm, tn);
                for (uint32_t kb = 0; kb < nkb; kb++) {
                    bar_wait(&empty[stage], phase ^ 1);                // MC: both CTAs of the cluster have released this slot
                    bar_expect_tx(&full[stage], STAGE_BYTES);
                    tma_load_2d(sA + stage * A_STAGE, &mapA, &full[stage], (int)(kb * BK), (int)(tm * BM));
                    if constexpr (MC)                                  // my half of the B tile goes to both CTAs
                        tma_load_2d_mc(sB + stage * B_STAGE + rank * (B_STAGE / 2), &mapB, &full[stage], (int)(kb * BK),
                                       (int)(tn * BN + rank * (BN / 2)), (uint16_t)3);
                    else
                        tma_load_2d(sB + stage * B_STAGE, &mapB, &full[stage], (int)(kb * BK), (int)(tn * BN));
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_i8(BM, BN);
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (uint32_t t = worker; t < nwork; t += n_workers) {
                bar_wait(&tempty[acc], acc_phase ^ 1);                 // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (uint32_t kb = 0; kb < nkb; kb++) {
                    bar_wait(&full[stage], phase);                     // TMA bytes have landed
                    tc_fence_after();
                    const uint64_t da = umma_desc_k_sw128(s32(sA + stage * A_STAGE));
                    const uint64_t db = umma_desc_k_sw128(s32(sB + stage * B_STAGE));
#pragma unroll
                    for (uint32_t k = 0; k < BK / 32; k++)             // UMMA_K = 32 int8 = 32 bytes = +2 in 16-byte units
                        tc_mma_i8(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    if constexpr (MC) tc_commit_mc(&empty[stage], (uint16_t)3);   // the slot is shared: release it in both CTAs
                    else tc_commit(&empty[stage]);                     // frees the smem slot when these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(&tfull[acc]);                                // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===== epilogue warps 2..5: TMEM lane quadrant = warp % 4 =====
        const uint32_t quad = warp & 3;
        uint32_t acc = 0, acc_phase = 0;
        const bool plain = outq.identity && outq.shift == 0;            // C's type holds the exact sum: no quantizer left
        for (uint32_t t = worker; t < nwork; t += n_workers) {
            uint32_t tm, tn; tile_coords(t, tm, tn);
            bar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint64_t row = (uint64_t)tm * BM + quad * 32 + lane;          // m index of this thread's TMEM lane
#pragma unroll 1
            for (uint32_t c0 = 0; c0 < (uint32_t)BN; c0 += 32) {
                uint32_t v[32];
                tc_ld32(tmem_base + ((quad * 32) << 16) + acc * BN + c0, v);
#pragma unroll
                for (int c = 0; c < 32; c++) {
                    const uint64_t col = (uint64_t)tn * BN + c0 + c;
                    const int32_t q = plain ? (int32_t)v[c] : conv_apply<int32_t, false>((int32_t)v[c], outq);   // converting assignment into C's type
                    if (row < P.M && col < P.N) {                                        // ragged shapes run on zero-padded operands
                        if (P.c_wb == 4) ((int32_t *)C)[col * P.M + row] = q;            // column-major: 32 lanes = 128 contiguous bytes
                        else             ((int64_t *)C)[col * P.M + row] = (int64_t)q;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) bar_arrive(&tempty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (MC) cluster_sync_all();          // no CTA leaves while its peer may still multicast into it
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
}

__global__ void __launch_bounds__(I8_THREADS, 1)
k_i8gemm(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, void *__restrict__ C,
         const __grid_constant__ I8Params P, const __grid_constant__ QStage outq)
{
    i8gemm_body<false>(mapA, mapB, C, P, outq);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(I8_THREADS, 1)
k_i8gemm_mc(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, void *__restrict__ C,
            const __grid_constant__ I8Params P, const __grid_constant__ QStage outq)
{
    i8gemm_body<true>(mapA, mapB, C, P, outq);
}

// ---------------------------------------------------------------------------------------------
// Pair variant: the two CTAs of a cluster (one TPC) execute ONE tcgen05.mma.cta_group::2 of M = 256, N = 256.  Each CTA
// stages only its own 128 rows of A and its own half of B (128 of the 256 columns), 32 KB per stage instead of 48, so six
// stages fit and every operand byte is read from shared memory once per pair instead of once per CTA.  Rank 0 issues the
// MMAs; both producers' TMA loads report to rank 0's `full` barrier (cta_group::2 form, peer bit cleared), the MMA
// commits are multicast to both CTAs (slot release and accumulator-ready), and both epilogues hand the accumulator back
// by arriving on rank 0's `tempty` barrier.  D rows 0..127 live in rank 0's TMEM, 128..255 in rank 1's.
// ---------------------------------------------------------------------------------------------
#ifndef QB_P2_STAGES
#define QB_P2_STAGES 6
#endif
static constexpr int P2_STAGES = QB_P2_STAGES;
static constexpr int P2_A_STAGE = BM * BK, P2_B_STAGE = (BN / 2) * BK;
static constexpr int P2_STAGE_BYTES = P2_A_STAGE + P2_B_STAGE;
static constexpr int P2_SMEM_BYTES = P2_STAGES * P2_STAGE_BYTES + 1024 + 256;

__device__ __forceinline__ void tma_load_2d_pair(void *dst, const CUtensorMap *map, uint64_t *leader_bar, int c0, int c1)
{
    const uint32_t bar = s32(leader_bar) & 0xFEFFFFFFu;              // clear the peer bit: the bytes are counted on rank 0's barrier
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(s32(dst)), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_mma_i8_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(s32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void bar_arrive_rank0(uint64_t *b)
{
    asm volatile(
        "{\n"
        ".reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, 0;\n"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
        "}\n" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void bar_wait_cluster(uint64_t *b, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAITC:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONEC;\n"
        "bra LAB_WAITC;\n"
        "DONEC:\n"
        "}\n" ::"r"(s32(b)), "r"(parity) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Fused operand narrowing (FUSE = true, round 2).  (An operand stored with its rows contiguous -- A as M x K, the
// reference's natural layout -- is transposed in registers by the same warps: see packer_warp.)  The reference keeps W <= 32 raws in int32 words; the tensor maps want
// int8.  Instead of two pack launches in front of the GEMM (HBM-bound, 2 x 48 us against 337 us of MMA time at 8192^3),
// eight extra "packer" warps per CTA narrow the words WHILE the tensor cores run: a warp claims a slice (8 operand rows
// x 512 k = 16 KiB of words) from a global counter, streams it (16-byte loads, 16-byte stores), and bumps the flag of its unit
// (256 rows x 512 k of the rows of one tile pair / the columns of one n-tile); the TMA producer polls the two flags
// of the k-chunk it is about to load (acquire + cross-proxy fence: generic-proxy stores, async-proxy reads).  Packers never
// wait on anything and all CTAs are co-resident (grid <= SM count), so every flag is reached.  Units are claimed in the
// order the rasterisation needs them: the first wave's panels k-chunk by k-chunk, then the remaining B panels, then the
// remaining A panels.
// ---------------------------------------------------------------------------------------------
static constexpr int PACK_WARPS = 8;
static constexpr int I8_THREADS_FUSED = I8_THREADS + 32 * PACK_WARPS;
static constexpr uint32_t PK_ROWS = 256;

struct I8Fuse {
    const int32_t *srcA, *srcB;   // raw words, K contiguous per operand row; nullptr: that operand is not packed by this kernel
    int8_t *dstA, *dstB;          // int8 staging the tensor maps sit on
    uint32_t *sync;               // [0] claim counter, [1 ..] flags of the A units (panel * nkc + kc), then of the B units
    uint32_t pk_k, slice_rows, slices;   // k extent of a unit (a multiple of BK), operand rows per claimed slice, slices per unit
    uint32_t nkc;                 // k-chunks of pk_k
    uint32_t panels_a, panels_b;  // units per k-chunk (0 when the operand is not packed here)
    uint32_t first_a, first_b;    // panels the first wave of tile pairs touches
    uint32_t group_a, group_b;    // panels that become live together afterwards: a row group of A, the n-tiles one more wave reaches
    uint32_t sleep_ns;            // experiment (QB200_I8_PACK_SLEEP): pause per slice once the first wave's panels are done
    uint32_t tr_a, tr_b;          // operand stored with its ROWS contiguous (A as M x K, B as N x K): the slice is transposed in registers
    uint32_t ld_a, ld_b;          // ... and its leading dimension (M resp. N)
    unsigned long long *stats;    // QB200_I8_STATS=1: [0] first CTA start, [1] last packer done, [2] ns the producers waited for flags, [3] last CTA done (globaltimer)
};
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

__device__ __forceinline__ void wait_unit(const uint32_t *flag, uint32_t target, unsigned long long *stats)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v >= target) { asm volatile("fence.proxy.async.global;" ::: "memory"); return; }
    const unsigned long long t0 = stats ? gtimer() : 0;
    for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v >= target) break;
        __nanosleep(32);
    }
    asm volatile("fence.proxy.async.global;" ::: "memory");        // the bytes were written through the generic proxy, TMA reads them through the async proxy
    if (stats) {
        const unsigned long long t1 = gtimer();
        atomicAdd(stats + 2, t1 - t0);
        const unsigned long long idx = atomicAdd(stats + 4, 1ull);
        if (idx < 4096) { stats[8 + 2 * idx] = t0; stats[9 + 2 * idx] = t1 - t0; }
    }
}

__device__ __forceinline__ void packer_warp(const I8Fuse &F, uint32_t K, int lane)
{
    const uint32_t units = (F.panels_a + F.panels_b) * F.nkc, total = units * F.slices;
    const uint32_t first = F.first_a + F.first_b, n0 = first * F.nkc, n1 = n0 + (F.panels_b - F.first_b) * F.nkc;
    uint32_t s = 0, s_next = 0;
    if (lane == 0) s = atomicAdd(F.sync, 1u);
    s = __shfl_sync(0xffffffffu, s, 0);
    while (s < total) {
        if (lane == 0) s_next = atomicAdd(F.sync, 1u);         // claimed now, looked at after this slice: the round trip hides behind the copy
        const uint32_t u = s / F.slices, slice = s - u * F.slices;
        // units in the order the rasterisation needs them, k-chunk-major inside every group of panels that becomes live together
        bool is_a; uint32_t panel, kc;
        if (u < n0) {                                   // the first wave's A and B panels
            kc = u / first; const uint32_t j = u - kc * first;
            is_a = j < F.first_a; panel = is_a ? j : j - F.first_a;
        } else {
            is_a = u >= n1;                             // then the B panels wave by wave, then the A panels row group by row group
            const uint32_t v = is_a ? u - n1 : u - n0, per = is_a ? F.group_a : F.group_b, lim = is_a ? F.panels_a : F.panels_b;
            const uint32_t g = v / (per * F.nkc), base = (is_a ? F.first_a : F.first_b) + g * per;
            const uint32_t cnt = (lim - base) < per ? (lim - base) : per, w = v - g * per * F.nkc;
            kc = w / cnt; panel = base + (w - kc * cnt);
        }
        if (F.sleep_ns && u >= n0) __nanosleep(F.sleep_ns);
        const uint32_t k0 = kc * F.pk_k, klen = (K - k0) < F.pk_k ? (K - k0) : F.pk_k, items = klen / 16;      // K % 128 == 0: items % 8 == 0
        if (is_a ? F.tr_a : F.tr_b) {
            // rows contiguous in memory: a lane owns one operand row, reads it k by k (the warp's 32 rows are 128 contiguous
            // bytes per k) and stores 32 narrowed k at a time (two 16-byte stores = one full sector of its K-major row).
            // The unit's slices are 8 blocks of 32 rows x (slices / 8) parts of the k-chunk.
            const uint32_t nsub = F.slices / 8, rb = slice % 8, ks = slice / 8, sub = klen / nsub;      // sub % 32 == 0 (host)
            const uint64_t ld = is_a ? F.ld_a : F.ld_b, row = (uint64_t)panel * PK_ROWS + rb * 32 + lane;
            const int32_t *src = (is_a ? F.srcA : F.srcB) + row + (uint64_t)(k0 + ks * sub) * ld;
            int8_t *dst = (is_a ? F.dstA : F.dstB) + row * K + k0 + ks * sub;
            for (uint32_t kk = 0; kk < sub; kk += 32) {
                int v[32];
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] = __ldcs(src + (uint64_t)(kk + j) * ld);
                uint4 *d = (uint4 *)(dst + kk);
                d[0] = make_uint4(pack4(v[0], v[1], v[2], v[3]), pack4(v[4], v[5], v[6], v[7]), pack4(v[8], v[9], v[10], v[11]), pack4(v[12], v[13], v[14], v[15]));
                d[1] = make_uint4(pack4(v[16], v[17], v[18], v[19]), pack4(v[20], v[21], v[22], v[23]), pack4(v[24], v[25], v[26], v[27]), pack4(v[28], v[29], v[30], v[31]));
            }
            asm volatile("fence.proxy.async.global;" ::: "memory");
            __threadfence();
            __syncwarp();
            if (lane == 0) {
                uint32_t *flag = F.sync + 1 + (is_a ? 0u : F.panels_a * F.nkc) + panel * F.nkc + kc;
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(flag) : "memory");
            }
            s = __shfl_sync(0xffffffffu, s_next, 0);
            continue;
        }
        const uint64_t row0 = (uint64_t)panel * PK_ROWS + slice * F.slice_rows;
        const int32_t *src = (is_a ? F.srcA : F.srcB) + row0 * K + k0;
        int8_t *dst = (is_a ? F.dstA : F.dstB) + row0 * K + k0;
        const uint32_t n_it = F.slice_rows * items;                // 16-word items in this slice, a multiple of 32
        // four items per lane and trip: sixteen 16-byte loads in flight before the first store
        for (uint32_t base = lane; base < n_it; base += 128) {
            int4 v[4][4];
            uint32_t r[4], t[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t i = base + 32 * j;
                r[j] = i / items; t[j] = i - r[j] * items;
                if (i < n_it) {
                    const int4 *p = (const int4 *)(src + (uint64_t)r[j] * K) + 4 * t[j];
                    v[j][0] = __ldcs(p); v[j][1] = __ldcs(p + 1); v[j][2] = __ldcs(p + 2); v[j][3] = __ldcs(p + 3);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (base + 32 * j < n_it)
                    *((uint4 *)(dst + (uint64_t)r[j] * K) + t[j]) =
                        make_uint4(pack4(v[j][0].x, v[j][0].y, v[j][0].z, v[j][0].w), pack4(v[j][1].x, v[j][1].y, v[j][1].z, v[j][1].w),
                                   pack4(v[j][2].x, v[j][2].y, v[j][2].z, v[j][2].w), pack4(v[j][3].x, v[j][3].y, v[j][3].z, v[j][3].w));
        }
        asm volatile("fence.proxy.async.global;" ::: "memory");
        __threadfence();
        __syncwarp();
        if (lane == 0) {
            uint32_t *flag = F.sync + 1 + (is_a ? 0u : F.panels_a * F.nkc) + panel * F.nkc + kc;
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(flag) : "memory");
        }
        s = __shfl_sync(0xffffffffu, s_next, 0);
    }
}

template <bool FUSE, bool FAN>
__device__ __forceinline__ void i8gemm_pair_body(const CUtensorMap &mapA, const CUtensorMap &mapB, void *__restrict__ C, const I8Params &P,
                                                 const QStage &outq, const I8Fuse &F)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char *sA = smem, *sB = smem + P2_STAGES * P2_A_STAGE;
    uint64_t *bars = (uint64_t *)(smem + P2_STAGES * P2_STAGE_BYTES);
    uint64_t *full = bars, *empty = bars + P2_STAGES, *tfull = bars + 2 * P2_STAGES, *tempty = bars + 2 * P2_STAGES + 2;
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * P2_STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t nkb = P.K / BK;
    const uint32_t rank = cluster_ctarank();
    const uint32_t worker = cluster_id_x(), n_workers = n_clusters_x();
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < P2_STAGES; s++) { bar_init(&full[s], 1); bar_init(&empty[s], 1); }
        for (int a = 0; a < 2; a++) { bar_init(&tfull[a], 1); bar_init(&tempty[a], 8); }       // 4 epilogue warps in each CTA
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const uint32_t rows = P.tiles_m / 2;                       // work item = a 256 x 256 tile pair
    const uint32_t nwork = rows * P.tiles_n;
    auto pair_coords = [&](uint32_t t, uint32_t &row, uint32_t &tn) {
        const uint32_t per_group = P.group_m * P.tiles_n;
        const uint32_t g = t / per_group, r = t - g * per_group;
        const uint32_t gm = (rows - g * P.group_m) < P.group_m ? (rows - g * P.group_m) : P.group_m;
        row = g * P.group_m + r % gm;
        tn = r / gm;
    };

    if (warp == 0) {
        // ===== TMA producer (both CTAs): my 128 rows of A, my 128 columns of B =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t t = worker; t < nwork; t += n_workers) {
                uint32_t row, tn; pair_coords(t, row, tn);
                for (uint32_t kb = 0; kb < nkb; kb++) {
                    if constexpr (FUSE) {
                        if (kb % (F.pk_k / BK) == 0) {                                    // a new k-chunk: has it been narrowed yet?
                            const uint32_t kc = kb / (F.pk_k / BK);
                            if (F.panels_a) wait_unit(F.sync + 1 + row * F.nkc + kc, F.slices, F.stats);
                            if (F.panels_b) wait_unit(F.sync + 1 + F.panels_a * F.nkc + tn * F.nkc + kc, F.slices, F.stats);
                        }
                    }
                    bar_wait(&empty[stage], phase ^ 1);                                   // released by the pair MMA's multicast commit
                    if (rank == 0) bar_expect_tx(&full[stage], 2 * P2_STAGE_BYTES);       // both CTAs' bytes land on rank 0's barrier
                    tma_load_2d_pair(sA + stage * P2_A_STAGE, &mapA, &full[stage], (int)(kb * BK), (int)((2 * row + rank) * BM));
                    tma_load_2d_pair(sB + stage * P2_B_STAGE, &mapB, &full[stage], (int)(kb * BK), (int)(tn * BN + rank * (BN / 2)));
                    if (++stage == P2_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread of rank 0 drives both tensor cores =====
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = umma_idesc_i8(2 * BM, BN);
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (uint32_t t = worker; t < nwork; t += n_workers) {
                bar_wait_cluster(&tempty[acc], acc_phase ^ 1);         // both epilogues have drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (uint32_t kb = 0; kb < nkb; kb++) {
                    bar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint64_t da = umma_desc_k_sw128(s32(sA + stage * P2_A_STAGE));
                    const uint64_t db = umma_desc_k_sw128(s32(sB + stage * P2_B_STAGE));
#pragma unroll
                    for (uint32_t k = 0; k < BK / 32; k++)
                        tc_mma_i8_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    tc_commit_pair(&empty[stage]);                     // slot free in both CTAs when these MMAs retire
                    if (++stage == P2_STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit_pair(&tfull[acc]);                           // accumulator complete in both CTAs
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (FUSE && warp >= 6) {
        if (F.stats && lane == 0) atomicMin(F.stats, gtimer());
        packer_warp(F, P.K, lane);                             // ===== packer warps 6..13 =====
        if (F.stats && lane == 0) atomicMax(F.stats + 1, gtimer());
    } else {
        // ===== epilogue warps 2..5 of both CTAs: my 128 rows of the pair =====
        const uint32_t quad = warp & 3;
        uint32_t acc = 0, acc_phase = 0;
        const bool plain = outq.identity && outq.shift == 0;
        for (uint32_t t = worker; t < nwork; t += n_workers) {
            uint32_t prow, tn; pair_coords(t, prow, tn);
            bar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint64_t row = (uint64_t)(2 * prow + rank) * BM + quad * 32 + lane;
#pragma unroll 1
            for (uint32_t c0 = 0; c0 < (uint32_t)BN; c0 += 32) {
                uint32_t v[32];
                tc_ld32(tmem_base + ((quad * 32) << 16) + acc * BN + c0, v);
#pragma unroll
                for (int c = 0; c < 32; c++) {
                    const uint64_t col = (uint64_t)tn * BN + c0 + c;
                    const int32_t q = plain ? (int32_t)v[c] : conv_apply<int32_t, false>((int32_t)v[c], outq);
                    if (row < P.M && col < P.N) {
                        if constexpr (FAN) {                                // GEMM -> all-gather: local and peer stores, 128 contiguous bytes per warp and column
                            const uint64_t e = col * P.ldc + P.row0 + row;
                            for (uint32_t o = 0; o < P.n_out; o++) ((int32_t *)P.outs[o])[e] = q;
                        } else {
                            if (P.c_wb == 4) ((int32_t *)C)[col * P.M + row] = q;
                            else             ((int64_t *)C)[col * P.M + row] = (int64_t)q;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) bar_arrive_rank0(&tempty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    if constexpr (FUSE) { if (F.stats && threadIdx.x == 0) atomicMax(F.stats + 3, gtimer()); }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(I8_THREADS, 1)
k_i8gemm_pair(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, void *__restrict__ C,
              const __grid_constant__ I8Params P, const __grid_constant__ QStage outq)
{
    i8gemm_pair_body<false, false>(mapA, mapB, C, P, outq, I8Fuse{});
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(I8_THREADS_FUSED, 1)
k_i8gemm_pair_fused(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, void *__restrict__ C,
                    const __grid_constant__ I8Params P, const __grid_constant__ QStage outq, const __grid_constant__ I8Fuse F)
{
    i8gemm_pair_body<true, false>(mapA, mapB, C, P, outq, F);
}

// the same two kernels with the result fanned out into the C of every device (qb200_gemul_panels_gathered): separate
// instantiations, so that the plain epilogue keeps its code
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(I8_THREADS, 1)
k_i8gemm_pair_fan(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, void *__restrict__ C,
                  const __grid_constant__ I8Params P, const __grid_constant__ QStage outq)
{
    i8gemm_pair_body<false, true>(mapA, mapB, C, P, outq, I8Fuse{});
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(I8_THREADS_FUSED, 1)
k_i8gemm_pair_fused_fan(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, void *__restrict__ C,
                        const __grid_constant__ I8Params P, const __grid_constant__ QStage outq, const __grid_constant__ I8Fuse F)
{
    i8gemm_pair_body<true, true>(mapA, mapB, C, P, outq, F);
}

// ---------------------------------------------------------------------------------------------
// pack: raw ABI words -> int8, K-major.  src is column-major R x Cn (R contiguous).
//   straight : dst[c][r] = src[r + c*R]     (K is already the contiguous index)
//   transpose: dst[r][c] = src[r + c*R]     (the contiguous index must become the row)
// ---------------------------------------------------------------------------------------------
template <typename W>
__global__ void k_pack_i8_straight(const W *__restrict__ src, int8_t *__restrict__ dst, uint64_t n)
{
    uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    for (; i + 3 < n; i += (uint64_t)gridDim.x * blockDim.x * 4) {
        char4 o;
        o.x = (signed char)src[i]; o.y = (signed char)src[i + 1]; o.z = (signed char)src[i + 2]; o.w = (signed char)src[i + 3];
        *(char4 *)(dst + i) = o;
    }
}
template <typename W>
__global__ void k_pack_i8_transpose(const W *__restrict__ src, int8_t *__restrict__ dst, uint32_t R, uint32_t Cn)
{
    __shared__ int8_t tile[64][65];
    const uint32_t r0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
    for (uint32_t i = threadIdx.y; i < 64; i += blockDim.y)          // coalesced read along r
        tile[i][threadIdx.x] = (int8_t)src[(uint64_t)(c0 + i) * R + r0 + threadIdx.x];
    __syncthreads();
    for (uint32_t i = threadIdx.y; i < 64; i += blockDim.y)          // coalesced write along c
        dst[(uint64_t)(r0 + i) * Cn + c0 + threadIdx.x] = tile[threadIdx.x][i];
}

// int32 words (the reference's storage for W <= 32), streamed at HBM rate: 16-byte loads, 128-byte row stores.
// straight: 4 independent 16-byte loads per thread, one 16-byte store; n % 16 == 0
__global__ void __launch_bounds__(256)
k_pack_i8_straight_w4(const int4 *__restrict__ src, uint4 *__restrict__ dst, uint64_t n16)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) {
        const int4 a = __ldcs(src + 4 * i), b = __ldcs(src + 4 * i + 1), c = __ldcs(src + 4 * i + 2), d = __ldcs(src + 4 * i + 3);
        dst[i] = make_uint4(pack4(a.x, a.y, a.z, a.w), pack4(b.x, b.y, b.z, b.w), pack4(c.x, c.y, c.z, c.w), pack4(d.x, d.y, d.z, d.w));
    }
}
// transpose: one CTA moves a 128 (r) x 128 (c) tile.  A lane is (lr, lc): 8 lanes along r read one full 128-byte line
// of a column as int4, 4 lane groups take 4 neighbouring column quads; each thread so holds a 4x4 block, packs the four
// columns of each of its rows into one word and parks it in shared memory (row stride 33 words: conflict-free both
// ways); rows then leave as 128 contiguous bytes.  R % 128 == 0, Cn % 128 == 0.
__global__ void __launch_bounds__(256)
k_pack_i8_transpose_w4(const int *__restrict__ src, uint32_t *__restrict__ dst, uint32_t R, uint32_t Cn)
{
    __shared__ uint32_t tile[128 * 33];
    const uint32_t r0 = blockIdx.x * 128, c0 = blockIdx.y * 128;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, lr = lane & 7, lc = lane >> 3;
#pragma unroll
    for (uint32_t it = 0; it < 4; ++it) {
        const uint32_t blk = warp * 4 + it;                    // 32 warp-iterations: 4 r-blocks of 32 x 8 c-blocks of 16
        const uint32_t rb = (blk & 3) * 32 + lr * 4, cb = (blk >> 2) * 16 + lc * 4;
        int4 v[4];
#pragma unroll
        for (uint32_t i = 0; i < 4; ++i) v[i] = __ldcs((const int4 *)(src + (uint64_t)(c0 + cb + i) * R + r0 + rb));
        uint32_t *t = tile + rb * 33 + (cb >> 2);
        t[0] = pack4(v[0].x, v[1].x, v[2].x, v[3].x);
        t[33] = pack4(v[0].y, v[1].y, v[2].y, v[3].y);
        t[66] = pack4(v[0].z, v[1].z, v[2].z, v[3].z);
        t[99] = pack4(v[0].w, v[1].w, v[2].w, v[3].w);
    }
    __syncthreads();
#pragma unroll
    for (uint32_t it = 0; it < 16; ++it) {
        const uint32_t r = warp * 16 + it;
        dst[((uint64_t)(r0 + r) * Cn + c0) / 4 + lane] = tile[r * 33 + lane];
    }
}

// Ragged shapes: dst is the zero-padded rows_p x k_p int8 operand (k contiguous); src holds rows x k words, either with
// k contiguous (transpose == 0: src[r * k + kk]) or with rows contiguous (transpose == 1: src[r + kk * rows]).
template <typename W>
__global__ void k_pack_i8_padded(const W *__restrict__ src, int8_t *__restrict__ dst, uint32_t rows, uint32_t k, uint32_t rows_p, uint32_t k_p, int transpose)
{
    __shared__ int8_t tile[32][33];
    const uint32_t r0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
    if (!transpose) {
        const uint32_t r = r0 + threadIdx.y, kk = k0 + threadIdx.x;
        if (r < rows_p && kk < k_p) dst[(uint64_t)r * k_p + kk] = (r < rows && kk < k) ? (int8_t)src[(uint64_t)r * k + kk] : (int8_t)0;
    } else {
        const uint32_t rs = r0 + threadIdx.x, ks = k0 + threadIdx.y;                 // read with rows contiguous
        tile[threadIdx.y][threadIdx.x] = (rs < rows && ks < k) ? (int8_t)src[(uint64_t)ks * rows + rs] : (int8_t)0;
        __syncthreads();
        const uint32_t r = r0 + threadIdx.y, kk = k0 + threadIdx.x;                  // write with k contiguous
        if (r < rows_p && kk < k_p) dst[(uint64_t)r * k_p + kk] = tile[threadIdx.x][threadIdx.y];
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// rows x K int8 matrix, K contiguous; box = 128 bytes of K x box_rows rows, 128B swizzle
static int make_map(CUtensorMap *m, const int8_t *base, uint64_t rows, uint64_t K, uint32_t box_rows)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return -1;
    cuuint64_t dims[2] = { K, rows };
    cuuint64_t strides[1] = { K };
    cuuint32_t box[2] = { (cuuint32_t)BK, box_rows };
    cuuint32_t estr[2] = { 1, 1 };
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -2;
}

static uint32_t round_up(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }
void i8gemm_padded_dims(uint32_t M, uint32_t N, uint32_t K, uint32_t *Mp, uint32_t *Np, uint32_t *Kp)
{
    *Mp = round_up(M, BM); *Np = round_up(N, BN); *Kp = round_up(K, BK);
}
// Tile multiples always qualify; ragged shapes run on zero-padded operands when the problem is large enough to be worth
// the tensor cores (>= 2^24 MACs) and the padding at most doubles the work.
bool i8gemm_shape_ok(uint32_t M, uint32_t N, uint32_t K)
{
    if (!M || !N || !K || M > (1u << 20) || N > (1u << 20) || K > (1u << 20)) return false;
    if (M % BM == 0 && N % BN == 0 && K % BK == 0) return true;
    uint32_t Mp, Np, Kp; i8gemm_padded_dims(M, N, K, &Mp, &Np, &Kp);
    const double real = (double)M * N * K, padded = (double)Mp * Np * Kp;
    return real >= (double)(1u << 24) && padded <= 2.0 * real;
}

template <typename W>
static void pack_operand(const void *src, int8_t *dst, uint32_t rows_contig, uint32_t cols, bool transpose, cudaStream_t s)
{
    // src is column-major rows_contig x cols
    if constexpr (sizeof(W) == 4) {
        const uint64_t n = (uint64_t)rows_contig * cols;
        if (!transpose && n % 16 == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0) {
            k_pack_i8_straight_w4<<<(unsigned)std::min<uint64_t>((n / 16 + 255) / 256, 148 * 16), 256, 0, s>>>((const int4 *)src, (uint4 *)dst, n / 16);
            count_launch(1);
            return;
        }
        if (transpose && rows_contig % 128 == 0 && cols % 128 == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 3) == 0) {
            k_pack_i8_transpose_w4<<<dim3(rows_contig / 128, cols / 128), 256, 0, s>>>((const int *)src, (uint32_t *)dst, rows_contig, cols);
            count_launch(1);
            return;
        }
    }
    if (!transpose) {
        const uint64_t n = (uint64_t)rows_contig * cols;
        k_pack_i8_straight<W><<<(unsigned)std::min<uint64_t>((n / 4 + 255) / 256, 148 * 32), 256, 0, s>>>((const W *)src, dst, n);
    } else {
        dim3 grid(rows_contig / 64, cols / 64), block(64, 4);
        k_pack_i8_transpose<W><<<grid, block, 0, s>>>((const W *)src, dst, rows_contig, cols);
    }
    count_launch(1);
}

template <typename W>
static void pack_padded(const void *src, int8_t *dst, uint32_t rows, uint32_t k, uint32_t rows_p, uint32_t k_p, bool transpose, cudaStream_t s)
{
    k_pack_i8_padded<W><<<dim3(k_p / 32, rows_p / 32), dim3(32, 32), 0, s>>>((const W *)src, dst, rows, k, rows_p, k_p, transpose ? 1 : 0);
    count_launch(1);
}

// An operand that already is what the tensor maps want -- int8 words, K contiguous (A stored K x M, i.e. trans_a; B stored
// K x N, i.e. !trans_b), 16-byte aligned base and row stride -- is read in place: the tensor map is built on the caller's
// buffer and no pack kernel runs.  Ragged extents need no padding either: TMA fills out-of-bounds box elements with zeros,
// and zeros add nothing to an exact dot product.
bool i8gemm_zero_copy(const void *ptr, int wb, bool k_major, uint32_t K)
{
    static int off = -1;                           // QB200_I8_NO_ZERO_COPY=1: always pack (experiments)
    if (off < 0) { const char *e = getenv("QB200_I8_NO_ZERO_COPY"); off = (e && e[0] == '1') ? 1 : 0; }
    return !off && wb == 1 && k_major && (((uintptr_t)ptr) & 15) == 0 && K % 16 == 0;
}

// One instance.  A, B: device pointers to raw ABI words (a_wb / b_wb bytes each), column-major as in the plan.
// Ap / Bp: int8 staging of Mp x Kp and Np x Kp bytes (i8gemm_padded_dims); may be NULL for an operand that is read in place.
// Which kernel variant runs: QB200_I8_VARIANT = 0 single-CTA kernel, 1 multicast clusters, 2 pair MMA (cta_group::2, default)
static int i8_variant()
{
    static int variant = -1;
    if (variant < 0) { const char *e = getenv("QB200_I8_VARIANT"); variant = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 2; }
    return variant;
}
// Words of the flag buffer the fused narrowing needs (0: this shape / these operands are packed by separate launches).
// An operand is narrowed inside the GEMM kernel when it holds int32 words with K contiguous on a tile-multiple shape.
// geometry of the fused narrowing: k extent of a unit and operand rows per claimed slice (QB200_I8_PK="k,rows": experiments)
static uint32_t g_pk_k = 0, g_pk_rows = 0;
static void pk_init()
{
    if (g_pk_k) return;
    uint32_t k = 1024, r = 8;
    if (const char *e = getenv("QB200_I8_PK")) { unsigned a = 0, b = 0; if (sscanf(e, "%u,%u", &a, &b) == 2 && a >= BK && a % BK == 0 && b && PK_ROWS % b == 0) { k = a; r = b; } }
    g_pk_rows = r; g_pk_k = k;
}
static uint32_t pk_k() { pk_init(); return g_pk_k; }
static uint32_t pk_rows() { pk_init(); return g_pk_rows; }
// rows contiguous (k_major == false): transposed in registers by the packers, needs slices in 8 row blocks x parts of >= 32 k
static bool fusable(const void *p, int wb, bool k_major, bool zero_copy)
{
    if (zero_copy || wb != 4) return false;
    if (k_major) return (((uintptr_t)p) & 15) == 0;
    const uint32_t slices = PK_ROWS / pk_rows();
    return (((uintptr_t)p) & 3) == 0 && slices % 8 == 0 && (BK / (slices / 8)) % 32 == 0 && !getenv("QB200_I8_NO_FUSE_T");
}
size_t i8gemm_sync_words(const void *A, const void *B, uint32_t M, uint32_t N, uint32_t K, bool trans_a, bool trans_b, int a_wb, int b_wb)
{
    static int off = -1;                           // QB200_I8_NO_FUSE=1: always pack with separate launches (experiments)
    if (off < 0) { const char *e = getenv("QB200_I8_NO_FUSE"); off = (e && e[0] == '1') ? 1 : 0; }
    if (off || i8_variant() != 2 || M % (2 * BM) || N % BN || K % BK) return 0;
    const bool fa = fusable(A, a_wb, trans_a, i8gemm_zero_copy(A, a_wb, trans_a, K)), fb = fusable(B, b_wb, !trans_b, i8gemm_zero_copy(B, b_wb, !trans_b, K));
    if (!fa && !fb) return 0;
    const uint64_t nkc = (K + pk_k() - 1) / pk_k(), units = ((fa ? M / PK_ROWS : 0) + (fb ? N / PK_ROWS : 0)) * nkc;
    if (units > (1u << 20)) return 0;
    return (size_t)units + 1;
}

int i8gemm_run(const void *A, const void *B, void *C, uint32_t M, uint32_t N, uint32_t K, bool trans_a, bool trans_b, int a_wb,
               int b_wb, int c_wb, const QStage &outq, int8_t *Ap, int8_t *Bp, uint32_t *sync, cudaStream_t s, const char **err, const I8FanOut *fan)
{
    uint32_t Mp, Np, Kp;
    i8gemm_padded_dims(M, N, K, &Mp, &Np, &Kp);
    const bool ragged = Mp != M || Np != N || Kp != K;
    bool zc_a = i8gemm_zero_copy(A, a_wb, trans_a, K), zc_b = i8gemm_zero_copy(B, b_wb, !trans_b, K);
    if ((!zc_a && !Ap) || (!zc_b && !Bp)) { *err = "missing int8 staging buffer"; return -1; }
    // A operand must become M rows x K (K contiguous).  Stored M x K column-major (M contiguous) unless trans_a (K x M, K contiguous).
    const bool a_tr = !trans_a;
    // B operand must become N rows x K.  Stored K x N column-major (K contiguous) unless trans_b (N x K, N contiguous).
    const bool b_tr = trans_b;
    // operands narrowed inside the GEMM kernel (k_i8gemm_pair_fused): no pack launch for them
    const size_t sync_words = sync ? i8gemm_sync_words(A, B, M, N, K, trans_a, trans_b, a_wb, b_wb) : 0;
    const bool fuse_a = sync_words && fusable(A, a_wb, trans_a, zc_a), fuse_b = sync_words && fusable(B, b_wb, !trans_b, zc_b);
    const bool real_zc_a = zc_a, real_zc_b = zc_b;
    if (fuse_a) zc_a = true;                       // skips the separate pack below; the tensor map still sits on the staging buffer
    if (fuse_b) zc_b = true;
    if (ragged) {
        if (!zc_a) switch (a_wb) {
        case 1: pack_padded<int8_t>(A, Ap, M, K, Mp, Kp, a_tr, s); break;
        case 2: pack_padded<int16_t>(A, Ap, M, K, Mp, Kp, a_tr, s); break;
        case 4: pack_padded<int32_t>(A, Ap, M, K, Mp, Kp, a_tr, s); break;
        default: pack_padded<int64_t>(A, Ap, M, K, Mp, Kp, a_tr, s); break;
        }
        if (!zc_b) switch (b_wb) {
        case 1: pack_padded<int8_t>(B, Bp, N, K, Np, Kp, b_tr, s); break;
        case 2: pack_padded<int16_t>(B, Bp, N, K, Np, Kp, b_tr, s); break;
        case 4: pack_padded<int32_t>(B, Bp, N, K, Np, Kp, b_tr, s); break;
        default: pack_padded<int64_t>(B, Bp, N, K, Np, Kp, b_tr, s); break;
        }
    } else {
        const uint32_t a_r = trans_a ? K : M, a_c = trans_a ? M : K;
        const uint32_t b_r = trans_b ? N : K, b_c = trans_b ? K : N;
        if (!zc_a) switch (a_wb) {
        case 1: pack_operand<int8_t>(A, Ap, a_r, a_c, a_tr, s); break;
        case 2: pack_operand<int16_t>(A, Ap, a_r, a_c, a_tr, s); break;
        case 4: pack_operand<int32_t>(A, Ap, a_r, a_c, a_tr, s); break;
        default: pack_operand<int64_t>(A, Ap, a_r, a_c, a_tr, s); break;
        }
        if (!zc_b) switch (b_wb) {
        case 1: pack_operand<int8_t>(B, Bp, b_r, b_c, b_tr, s); break;
        case 2: pack_operand<int16_t>(B, Bp, b_r, b_c, b_tr, s); break;
        case 4: pack_operand<int32_t>(B, Bp, b_r, b_c, b_tr, s); break;
        default: pack_operand<int64_t>(B, Bp, b_r, b_c, b_tr, s); break;
        }
    }
    zc_a = real_zc_a; zc_b = real_zc_b;
    const int variant = i8_variant();
    const bool pair = variant == 2 && (Mp / BM) % 2 == 0;
    const bool mc = (variant == 1 || pair) && (Mp / BM) % 2 == 0;
    CUtensorMap mapA, mapB;
    // in place: the real extents (out-of-bounds box elements read as zero); staged: the zero-padded copy
    if ((zc_a ? make_map(&mapA, (const int8_t *)A, M, K, BM) : make_map(&mapA, Ap, Mp, Kp, BM)) ||
        (zc_b ? make_map(&mapB, (const int8_t *)B, N, K, mc ? BN / 2 : BN) : make_map(&mapB, Bp, Np, Kp, mc ? BN / 2 : BN))) { *err = "cuTensorMapEncodeTiled failed"; return -1; }
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    static bool attr_set[64] = { false };          // function attributes are per device
    if (!attr_set[dev & 63]) {
        if (cudaFuncSetAttribute(k_i8gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(k_i8gemm_mc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(k_i8gemm_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(k_i8gemm_pair_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(k_i8gemm_pair_fan, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(k_i8gemm_pair_fused_fan, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM_BYTES) != cudaSuccess) { *err = "cannot raise dynamic smem"; return -1; }
        attr_set[dev & 63] = true;
    }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    I8Params P;
    P.M = M; P.N = N; P.K = Kp; P.tiles_m = Mp / BM; P.tiles_n = Np / BN; P.c_wb = c_wb; P.group_m = 16;
    P.n_out = 0; P.row0 = 0; P.ldc = M;
    for (int o = 0; o < 8; o++) P.outs[o] = nullptr;
    if (fan) {
        if (!pair || c_wb != 4 || fan->n == 0 || fan->n > 8) { *err = "the fused all-gather needs the pair kernel (M a multiple of 256), int32 results and at most 8 destinations"; return -1; }
        P.n_out = fan->n; P.row0 = fan->row0; P.ldc = fan->ldc;
        for (uint32_t o = 0; o < fan->n; o++) P.outs[o] = fan->out[o];
    }
    const uint32_t ntiles = P.tiles_m * P.tiles_n;
    if (mc) {
        P.group_m = 8;                                                     // 8 tile pairs = 16 m-tiles per group
        const uint32_t npairs = ntiles / 2, clusters = npairs < (uint32_t)(sms / 2) ? npairs : (uint32_t)(sms / 2);
        if (pair && (fuse_a || fuse_b)) {
            I8Fuse F;
            F.srcA = fuse_a ? (const int32_t *)A : nullptr; F.srcB = fuse_b ? (const int32_t *)B : nullptr;
            F.dstA = Ap; F.dstB = Bp; F.sync = sync;
            F.pk_k = pk_k(); F.slice_rows = pk_rows(); F.slices = PK_ROWS / F.slice_rows;
            F.nkc = (K + F.pk_k - 1) / F.pk_k;
            F.panels_a = fuse_a ? M / PK_ROWS : 0; F.panels_b = fuse_b ? N / PK_ROWS : 0;
            const uint32_t rows = P.tiles_m / 2, gm = rows < P.group_m ? rows : P.group_m;
            const uint32_t wave_n = (clusters + gm - 1) / gm + 1;                  // n-tiles the first wave of tile pairs reaches
            F.first_a = fuse_a ? gm : 0;
            F.first_b = fuse_b ? (wave_n < P.tiles_n ? wave_n : P.tiles_n) : 0;
            F.group_a = gm; F.group_b = wave_n > 1 ? wave_n - 1 : 1;
            F.tr_a = a_tr ? 1u : 0u; F.tr_b = b_tr ? 1u : 0u; F.ld_a = M; F.ld_b = N;
            static int packers = 0, sleep_ns = -1;  // experiments: QB200_I8_PACKERS = packer warps per CTA (1..8), QB200_I8_PACK_SLEEP = ns per slice after the first wave
            if (!packers) { const char *e = getenv("QB200_I8_PACKERS"); const int v = e ? atoi(e) : 0; packers = (v >= 1 && v <= PACK_WARPS) ? v : PACK_WARPS; }
            if (sleep_ns < 0) { const char *e = getenv("QB200_I8_PACK_SLEEP"); const int v = e ? atoi(e) : 0; sleep_ns = v > 0 ? v : 0; }
            F.sleep_ns = (uint32_t)sleep_ns;
            const unsigned fused_threads = I8_THREADS + 32 * packers;
            if (cudaMemsetAsync(sync, 0, sync_words * sizeof(uint32_t), s) != cudaSuccess) { *err = "cudaMemsetAsync failed"; return -1; }
            static int want_stats = -1;            // QB200_I8_STATS=1: per-call timeline of the fused kernel on stderr (synchronises; experiments only)
            if (want_stats < 0) { const char *e = getenv("QB200_I8_STATS"); want_stats = (e && e[0] == '1') ? 1 : 0; }
            F.stats = nullptr;
            if (want_stats) {
                const unsigned long long init[8] = { ~0ull, 0, 0, 0, 0, 0, 0, 0 };
                if (cudaMalloc(&F.stats, (8 + 2 * 4096) * sizeof(unsigned long long)) != cudaSuccess) { *err = "cudaMalloc failed"; return -1; }
                cudaMemcpyAsync(F.stats, init, sizeof init, cudaMemcpyHostToDevice, s);
            }
            if (P.n_out) k_i8gemm_pair_fused_fan<<<clusters * 2, fused_threads, P2_SMEM_BYTES, s>>>(mapA, mapB, C, P, outq, F);
            else         k_i8gemm_pair_fused<<<clusters * 2, fused_threads, P2_SMEM_BYTES, s>>>(mapA, mapB, C, P, outq, F);
            if (want_stats) {
                static unsigned long long st[8 + 2 * 4096];
                cudaMemcpyAsync(st, F.stats, sizeof st, cudaMemcpyDeviceToHost, s);
                cudaStreamSynchronize(s);
                cudaFree(F.stats);
                {   // histogram of the waits over the kernel's timeline (20 us bins): when do the producers starve?
                    double bins[64] = { 0 }; unsigned cnt[64] = { 0 };
                    const unsigned long long n = st[4] < 4096 ? st[4] : 4096;
                    for (unsigned long long i = 0; i < n; i++) {
                        unsigned b = (unsigned)((st[8 + 2 * i] - st[0]) / 20000); if (b > 63) b = 63;
                        bins[b] += st[9 + 2 * i] * 1e-3; cnt[b]++;
                    }
                    for (int b = 0; b < 64; b++) if (cnt[b]) fprintf(stderr, "  waits starting in [%d, %d) us: %u, %.1f us in total\n", b * 20, b * 20 + 20, cnt[b], bins[b]);
                }
                fprintf(stderr, "k_i8gemm_pair_fused: packers done after %.1f us, kernel %.1f us, producers waited %.1f us in total (%u producers)\n",
                        (st[1] - st[0]) * 1e-3, (st[3] - st[0]) * 1e-3, st[2] * 1e-3, clusters * 2);
            }
        }
        else if (pair && P.n_out) k_i8gemm_pair_fan<<<clusters * 2, I8_THREADS, P2_SMEM_BYTES, s>>>(mapA, mapB, C, P, outq);
        else if (pair) k_i8gemm_pair<<<clusters * 2, I8_THREADS, P2_SMEM_BYTES, s>>>(mapA, mapB, C, P, outq);
        else      k_i8gemm_mc<<<clusters * 2, I8_THREADS, SMEM_BYTES, s>>>(mapA, mapB, C, P, outq);
    } else {
        const unsigned grid = ntiles < (uint32_t)sms ? ntiles : (unsigned)sms;
        k_i8gemm<<<grid, I8_THREADS, SMEM_BYTES, s>>>(mapA, mapB, C, P, outq);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return -1; }
    count_launch(1);
    return 0;
}

}  // namespace qb
