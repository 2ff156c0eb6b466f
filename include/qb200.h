/*
 * qb200.h — plain-C ABI of the B200-native QuBLAS hot path (libqb200.so).
 *
 * The reference (pikapuma/QuBLAS, header-only C++23) has no FFI seam: its boundary is
 * the template API itself.  This ABI is what a binding of that API lowers to: every
 * `Qu<...>` type becomes a POD `qb_qdesc`, every per-level `TypeList` a `qb_levels`,
 * every tensor a buffer of raw two's-complement words in the reference's own memory
 * layout (`ArbiInt<N>::data`, QuBLAS.h:353 — int32 for N<=32, int64 for N<=64;
 * tensors column-major, QuBLAS.h:2680-2692; complex elements AoS {real, imag},
 * QuBLAS.h:2501-2617).
 *
 * Reference interface each entry point replaces (file:line in /root/reference):
 *   qb200_reduce_batched   <- Qreduce<Args...>(tensor / scalars...)   include/QuBLAS.h:4903-5018
 *   qb200_gemul_batched    <- Qgemul<QgemulAddArgs<>,QgemulMulArgs<>,QgemulTransposedA<>>(C,A,B)
 *                             readme.md:80-87 (documented only; composed from Qmul :3980-3985,
 *                             Qreduce :5014-5018 and the converting ctor :2398-2411)
 *   qb200_cgemul_batched   <- the same call with Qcomplex elements and
 *                             MulArgs = TFComplexMul<abT..BCT>       include/QuBLAS.h:3453-3535,
 *                             accumulated with complex Qadd           include/QuBLAS.h:3549-3568
 *   qb200_tfcmul_batched   <- tensor Qmul<TFComplexMul<...>>(x, y) / Qmul<BasicComplexMul<...>>(x, y)
 *                             include/QuBLAS.h:3783-3811, 4080-4085, 3426-3446
 *   qb200_convert_batched  <- element-wise converting assignment      include/QuBLAS.h:2398-2411, 2759-2830
 *   qb200_poly_batched     <- ANUS::Qpoly / Qapprox over a tensor         include/QuBLAS.h:4829-4897
 *
 * All functions return 0 on success and a negative QB_E* code on failure; the message is
 * available from qb200_last_error() (thread-local).  Nothing throws.  There is no CPU
 * fallback: without a CUDA device every compute entry point fails with QB_ENODEVICE.
 */
#ifndef QB200_H
#define QB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QB200_ABI_VERSION 2
#define QB_MAX_LEVELS 64

/* QuMode ids == the reference's ::value constants (QuBLAS.h:1988-1998). */
enum qb_qumode {
    QB_RND_POS_INF = 0, QB_RND_NEG_INF = 1, QB_RND_ZERO = 2, QB_RND_INF = 3, QB_RND_CONV = 4,
    QB_TRN_TCPL = 5, QB_TRN_SMGN = 6
};
/* OfMode ids == the reference's ::value constants (QuBLAS.h:2212-2223). */
enum qb_ofmode { QB_SAT_TCPL = 0, QB_SAT_ZERO = 1, QB_SAT_SMGN = 2, QB_WRP_TCPL = 3 };

enum qb_status {
    QB_OK = 0,
    QB_EINVAL = -1,      /* malformed plan / null pointer                                   */
    QB_EDOMAIN = -2,     /* plan is outside the bit-exact parity domain     */
    QB_ENODEVICE = -3,   /* no CUDA device / driver                                         */
    QB_ECUDA = -4,       /* a CUDA call failed                                              */
    QB_ENOMEM = -5
};

/* Where the caller's buffers live. */
enum qb_memkind { QB_MEM_HOST = 0, QB_MEM_DEVICE = 1 };

/* Tree form for non-power-of-two counts (QuBLAS.h:4924-4957 vs :4960-4990). */
enum qb_treeform { QB_TREE_VECTOR = 0, QB_TREE_VARIADIC = 1 };

/* One fixed-point type: Qu<intBits<i>, fracBits<f>, isSigned<s>, QuMode<q>, OfMode<o>>
 * (QuBLAS.h:2368-2385).  Stored width is always 1 + int_bits + frac_bits.
 * word_bytes: storage word of buffers holding this type: 0 = the reference's natural word
 * (4 if width<=32 else 8); 1/2/4/8 select a packed word (sign-extended on load). */
typedef struct qb_qdesc {
    int16_t int_bits;
    int16_t frac_bits;
    uint8_t is_signed;
    uint8_t qu_mode;   /* enum qb_qumode */
    uint8_t of_mode;   /* enum qb_ofmode */
    uint8_t word_bytes;
} qb_qdesc;

/* Per-level accumulation types: level l uses lv[min(l, n-1)] (QuBLAS.h:4913).
 * n == 0 means "no AddArgs": every level uses the default AddMerger of its operands
 * (QuBLAS.h:3125-3140). */
typedef struct qb_levels {
    uint32_t n;
    qb_qdesc lv[QB_MAX_LEVELS];
} qb_levels;

/* Qreduce over `batch` independent vectors of `len` elements each. */
typedef struct qb_reduce_plan {
    qb_qdesc  elem;
    qb_levels levels;
    uint64_t  len;
    uint32_t  tree_form;      /* enum qb_treeform */
    uint32_t  reserved;
} qb_reduce_plan;

/* Real Qgemul: C[i,j] = conv_c( tree_k( Qmul<mul>(opA(i,k), opB(k,j)) ) ).
 * A is M x K (K x M when trans_a), B is K x N (N x K when trans_b), C is M x N; all column-major.
 * stride_* are in elements between consecutive instances; 0 for A/B = shared by all instances. */
typedef struct qb_gemul_plan {
    qb_qdesc  a, b, mul, c;
    qb_levels add;
    uint32_t  M, N, K;
    uint8_t   trans_a, trans_b, tree_form, reserved;
    int64_t   stride_a, stride_b, stride_c;
} qb_gemul_plan;

/* Resolved types of the TFComplexMul dataflow (QuBLAS.h:3510-3535), x = a+bi, y = c+di:
 *   ab = a+b, cd = c+d, ba = b-a, A = ab*c (abc), B = cd*b (bad), C = ba*d (cdb),
 *   re = A-B (AB), im = B-C (BC).
 * The reference's routing quirks (badT governs B, cdbT governs C, baT is ignored) are applied
 * by the host lowering, not here: these are the types each node really has. */
typedef struct qb_tfc_types {
    qb_qdesc ab, cd, ba, abc, bad, cdb, AB, BC;
} qb_tfc_types;

/* Resolved types of the BasicComplexMul dataflow (QuBLAS.h:3426-3446), x = a+bi, y = c+di:
 *   ac, bd, ad, bc = the four products; re = ac - bd (acbd), im = ad + bc (adbc). */
typedef struct qb_bcm_types {
    qb_qdesc ac, bd, ad, bc, acbd, adbc;
} qb_bcm_types;

/* QB_CPROD_REAL_A / QB_CPROD_REAL_B: one operand is a REAL tensor (one word per element, typed by
 * a_re resp. b_re; the matching *_im descriptor is ignored) and scales both components of the
 * complex one (real (x) complex Qmul, QuBLAS.h:3600-3637).  Product types live in .bcm:
 *   REAL_A  x = a,      y = c + di :  re = Qmul<bcm.ac>(a, c), im = Qmul<bcm.ad>(a, d)
 *   REAL_B  x = a + bi, y = c      :  re = Qmul<bcm.ac>(a, c), im = Qmul<bcm.bc>(b, c)      */
enum qb_cproduct { QB_CPROD_TF = 0, QB_CPROD_BASIC = 1, QB_CPROD_REAL_A = 2, QB_CPROD_REAL_B = 3 };

/* Complex Qgemul with TFComplexMul (product == QB_CPROD_TF, uses .tf) or BasicComplexMul
 * (product == QB_CPROD_BASIC, uses .bcm) products and component-wise tree accumulation. */
typedef struct qb_cgemul_plan {
    qb_qdesc     a_re, a_im, b_re, b_im;
    qb_tfc_types tf;
    qb_bcm_types bcm;
    qb_levels    add_re, add_im;
    qb_qdesc     c_re, c_im;
    uint32_t     M, N, K;
    uint8_t      trans_a, trans_b, tree_form, product;   /* product: enum qb_cproduct */
    int64_t      stride_a, stride_b, stride_c;   /* in complex elements */
} qb_cgemul_plan;

/* Element-wise complex product of two complex vectors: TFComplexMul (output type {AB, BC}) or
 * BasicComplexMul (output type {acbd, adbc}). */
typedef struct qb_tfcmul_plan {
    qb_qdesc     x_re, x_im, y_re, y_im;
    qb_tfc_types tf;
    qb_bcm_types bcm;
    uint64_t     len;
    uint32_t     product;     /* enum qb_cproduct */
    uint32_t     reserved;
} qb_tfcmul_plan;

/* Element-wise Qmul / Qadd / Qsub of two real tensors (the reference's lazy Mul/Add/SubExpression evaluated
 * eagerly, include/QuBLAS.h:3741-3811, 4080-4156): out[i] = op<to>(a[i], b[i]); `to` = resolved merger type. */
enum qb_ewop { QB_EW_MUL = 0, QB_EW_ADD = 1, QB_EW_SUB = 2 };
typedef struct qb_ew_plan {
    qb_qdesc a, b, to;
    uint64_t len;
    uint32_t op;          /* enum qb_ewop */
    uint32_t reserved;
} qb_ew_plan;

/* Element-wise conversion from -> to (converting ctor, target's modes). */
typedef struct qb_convert_plan {
    qb_qdesc from, to;
    uint64_t len;
} qb_convert_plan;

/* ANUS::Qpoly / Segment / Qapprox on tensors (include/QuBLAS.h:4829-4897; the reference evaluates one scalar per call).
 * A segment is Qpoly<a0, a1, ...>(x) = a0 + x*(a1 + x*(...)) by Horner's rule: at step i the product Qmul<Ai>(x, inner)
 * and the sum Qadd<Ai>(ai, product) are quantized to Ai = the type of coefficient ai (QuBLAS.h:4846-4853).
 *   approx == 0: nseg must be 1; out[e] = Qpoly(x[e]), result type = coef[0].type (a single coefficient is returned as is).
 *   approx == 1: Qapprox<Segment<bp0, ...>, ...>: the first segment whose breakpoint exceeds x.toDouble(), the last one
 *                otherwise; the value is converted back into x's own type (QuBLAS.h:4866-4888).  breakpoint of the last
 *                segment is unused.  x wider than 53 bits is refused (toDouble would round). */
#define QB_POLY_MAX_COEF 8
#define QB_POLY_MAX_SEG  8
typedef struct qb_poly_coef { qb_qdesc type; int64_t raw; } qb_poly_coef;
typedef struct qb_poly_segment {
    double       breakpoint;
    uint32_t     ncoef, reserved;
    qb_poly_coef coef[QB_POLY_MAX_COEF];
} qb_poly_segment;
typedef struct qb_poly_plan {
    qb_qdesc        x;
    uint32_t        nseg, approx;
    qb_poly_segment seg[QB_POLY_MAX_SEG];
    uint64_t        len;
} qb_poly_plan;

/* What the library decided for a plan (for tests, benchmarks and DESIGN.md tables). */
typedef struct qb_plan_info {
    int32_t arith_bits;        /* 32 or 64: register arithmetic class of the kernel            */
    int32_t max_inter_bits;    /* widest intermediate (two's complement bits) the plan produces */
    int32_t lossless;          /* 1 if every quantizer on the accumulation path is an identity  */
    int32_t tensor_core;       /* 1 if the tcgen05 kind::i8 path is selected                    */
    int32_t specialized;       /* 1 if a compile-time-specialised kernel is used                */
    int32_t out_word_bytes;    /* storage word of the result buffer                             */
    qb_qdesc out_type;         /* type of the result (reduce: last level reached)               */
    char    kernel[96];        /* kernel variant name                                           */
} qb_plan_info;

/* exec.device = QB_DEVICE_ALL (batched calls on HOST buffers): the batch is cut into contiguous shards, one per visible
 * GPU, which run concurrently (one host thread, stream set and staging pipeline per device); no inter-GPU traffic. */
#define QB_DEVICE_ALL (-2)

/* Execution context: where buffers live, which device, which stream (cudaStream_t or NULL). */
typedef struct qb_exec {
    int32_t  mem_kind;   /* enum qb_memkind */
    int32_t  device;     /* CUDA device ordinal; -1 = current device; QB_DEVICE_ALL = every visible device */
    void    *stream;     /* cudaStream_t; NULL = default stream. Host-memory calls synchronise. */
    uint32_t flags;      /* QB_F_*                                                             */
    uint32_t reserved;
} qb_exec;

#define QB_F_NO_TENSOR_CORE 0x1u   /* force the integer-ALU path even when kind::i8 is legal  */
#define QB_F_NO_SPECIALIZE  0x2u   /* force the runtime-descriptor kernels                    */

int         qb200_abi_version(void);
const char *qb200_last_error(void);
int         qb200_device_count(void);
/* Number of kernels launched by this library (process-wide, all threads) since the last reset. */
uint64_t    qb200_launch_count(void);
void        qb200_launch_count_reset(void);

/* Plan validation + kernel selection without running anything (no device needed). */
int qb200_reduce_plan_info(const qb_reduce_plan *plan, qb_plan_info *info);
int qb200_gemul_plan_info(const qb_gemul_plan *plan, qb_plan_info *info);
int qb200_cgemul_plan_info(const qb_cgemul_plan *plan, qb_plan_info *info);

/* x: batch*len words (elem word).  out: batch words of the result type.
 * partials (optional, may be NULL): per-vector level-`partial_level` nodes,
 * ceil(len / 2^partial_level) words each, emitted for depth-wise parity checks. */
int qb200_reduce_batched(const qb_reduce_plan *plan, const void *x, void *out,
                         void *partials, uint32_t partial_level,
                         uint64_t batch, const qb_exec *exec);

int qb200_gemul_batched(const qb_gemul_plan *plan, const void *A, const void *B, void *C,
                        uint64_t batch, const qb_exec *exec);

/* Complex buffers are AoS {re, im}; both parts use the wider of the two natural words. */
int qb200_cgemul_batched(const qb_cgemul_plan *plan, const void *A, const void *B, void *C,
                         uint64_t batch, const qb_exec *exec);

int qb200_tfcmul_batched(const qb_tfcmul_plan *plan, const void *x, const void *y, void *out,
                         uint64_t batch, const qb_exec *exec);

int qb200_convert_batched(const qb_convert_plan *plan, const void *in, void *out,
                          uint64_t batch, const qb_exec *exec);

int qb200_elementwise_batched(const qb_ew_plan *plan, const void *a, const void *b, void *out,
                              uint64_t batch, const qb_exec *exec);

/* ---- non-linearities around the GEMM ----
 * qb200_poly_batched       <- ANUS::Qpoly<a0, ...>(x) / ANUS::Qapprox<Segment<...>...>(x), element by element over a tensor
 * qb200_gemul_poly_batched <- the same applied to every element of a Qgemul result before it is stored (fused epilogue):
 *                             gemul->c is the type Qgemul converts into and must equal poly->x; C holds words of the
 *                             polynomial's result type.  poly->len is ignored.
 * qb200_table_batched      <- ANUS::Qtable<func>(x) of readme.md:66-78 (absent from the header): a lookup table indexed by
 *                             x's raw value; table[r - raw_min(x)] is the raw result word (out type) for raw input r, built by
 *                             the host binding as out_type(func(x.toDouble())).  x wider than 20 bits is refused. */
int qb200_poly_plan_info(const qb_poly_plan *plan, qb_plan_info *info);
int qb200_poly_batched(const qb_poly_plan *plan, const void *x, void *out, uint64_t batch, const qb_exec *exec);
int qb200_gemul_poly_batched(const qb_gemul_plan *gemul, const qb_poly_plan *poly, const void *A, const void *B, void *C,
                             uint64_t batch, const qb_exec *exec);
/* The polynomial as a PROLOGUE: applied to the elements of operand A and / or B before the products
 * (Qgemul(C, ANUS::Qapprox<...>(A), B) evaluated without materialising the operand).  poly_a / poly_b may be NULL; a
 * polynomial's result type must equal the plan's operand type, its x words have the operand's word size.  One launch with the
 * run-time specialised kernels (the polynomial runs as the operand is staged), otherwise the polynomial runs into a scratch copy. */
int qb200_gemul_prologue_batched(const qb_gemul_plan *gemul, const qb_poly_plan *poly_a, const qb_poly_plan *poly_b, const void *A, const void *B,
                                 void *C, uint64_t batch, const qb_exec *exec);
int qb200_table_batched(const qb_qdesc *x, const qb_qdesc *out_type, const void *table, const void *in, void *out, uint64_t n,
                        const qb_exec *exec);

/* ---- the steps on either side of the path in an accuracy sweep ----
 * qb200_quantize_f64  <- Qu_s(double)  include/QuBLAS.h:2387-2393 (exact value, the type's QuMode, then its OfMode)
 * qb200_dequantize_f64 <- toDouble()   include/QuBLAS.h:2413-2416
 * qb200_error_stats: sums for SQNR (10 log10(sum_sq_ref / sum_sq_err)) and the max abs error of a fixed-point
 * buffer against a float64 reference, reduced on the device in a fixed order (deterministic). */
typedef struct qb_err_stats {
    double   sum_sq_err, sum_sq_ref, max_abs_err;
    uint64_t n;
} qb_err_stats;
int qb200_quantize_f64(const qb_qdesc *to, const double *in, void *out, uint64_t n, const qb_exec *exec);
int qb200_dequantize_f64(const qb_qdesc *from, const void *in, double *out, uint64_t n, const qb_exec *exec);
int qb200_error_stats(const qb_qdesc *type, const void *words, const double *ref, uint64_t n, qb_err_stats *stats, const qb_exec *exec);

/* float64 GEMM of the dequantized operands of a real Qgemul plan (same shapes, transposes and batch strides; C64 is
 * batch x M x N doubles, column-major): the yardstick qb200_error_stats compares a quantized result with. */
int qb200_gemul_f64_batched(const qb_gemul_plan *plan, const void *A, const void *B, double *C64, uint64_t batch, const qb_exec *exec);
/* The two steps above in one pass over a batch: error statistics of the quantized results C (the plan's C type) against the
 * float64 GEMM of the dequantized operands, without materialising the float64 matrices (the accuracy sweep of a Monte-Carlo
 * run: Qu -> double on either side of the path, QuBLAS.h:2413-2416).  Same reference values as qb200_gemul_f64_batched; the
 * sums are reduced in a fixed order (reproducible), which differs from qb200_error_stats' order in the last bits. */
int qb200_gemul_error_batched(const qb_gemul_plan *plan, const void *A, const void *B, const void *C, uint64_t batch,
                              qb_err_stats *stats, const qb_exec *exec);

/* Memory helpers for the host bindings.  Pinned (page-locked) host buffers make the copies of QB_MEM_HOST calls truly
 * asynchronous, so that the library's internal H2D / kernel / D2H pipeline overlaps; device buffers keep a whole
 * Monte-Carlo sweep (synth -> Qgemul -> error statistics) on the GPU.  qb200_memcpy is a synchronous copy in any direction. */
enum qb_allocflags { QB_ALLOC_DEFAULT = 0, QB_ALLOC_WRITE_COMBINED = 1 /* host: fast for the GPU to read, slow for the CPU to read */ };
int qb200_host_alloc(void **ptr, size_t bytes, uint32_t flags);
int qb200_host_free(void *ptr);
int qb200_device_alloc(void **ptr, size_t bytes, int device);
int qb200_device_free(void *ptr, int device);
int qb200_memcpy(void *dst, const void *src, size_t bytes, int device);

/* Deterministic synthetic raw words, uniform over [lo, hi], counter-based:
 * word[i] = lo + floor(mix64(mix64(seed * C + stream_id) ^ (first_index + i)) * (hi - lo + 1) / 2^64)
 * (splitmix64 finaliser, multiply-high range reduction; oracle/qb_oracle.c:qbo_synth_word is the twin).
 * Same function on host and device; buffers are device memory when
 * exec->mem_kind == QB_MEM_DEVICE. */
int qb200_synth_fill(void *dst, uint64_t count, uint32_t word_bytes, int64_t lo, int64_t hi,
                     uint64_t seed, uint64_t stream_id, uint64_t first_index, const qb_exec *exec);

/* Diagnostic: measured INT32 issue peak of the current device in tera thread-operations per second
 * (dependent-chain-free mix of IMAD on the FMA pipe and IADD3/LOP3 on the ALU pipe, all SMs busy).
 * This is the roofline denominator for the integer-ALU-bound kernels (MEASURED_PEAKS.json has no
 * integer figure).  fma_share in [0,1] = fraction of IMAD instructions in the mix. */
int qb200_measure_int_peak(double fma_share, double *tera_ops_per_s, const qb_exec *exec);

/* Diagnostic: run-time specialisation (NVRTC) of a Qgemul plan, compile step only (no device needed).
 * Returns 0 if a cubin was produced, 1 if the plan does not fit the specialised kernel template,
 * -1 if NVRTC is unavailable or compilation failed (message in `log`). */
int qb200_jit_check_gemul(const qb_gemul_plan *plan, char *log, uint64_t log_cap);
/* The same check for the other run-time specialised kernel templates (no device needed).
 * what: 1 = qb_cgemul_plan, 2 = qb_reduce_plan, 3 = qb_tfcmul_plan, 4 = qb_ew_plan, 5 = qb_convert_plan, 6 = qb_poly_plan,
 * 7 = struct { qb_gemul_plan; qb_poly_plan; } (Qgemul with the polynomial fused into its store), 8 = the same pair with the
 * polynomial fused into the loads of both operands.
 * Returns 0 = cubin produced, 1 = no template applies to this plan, -1 = NVRTC failed (log filled), < -1 = plan error. */
int qb200_jit_check(int what, const void *plan, char *log, uint64_t log_cap);

/* Diagnostic / parity evidence: exhaustive sweep of the quantizers that the specialised kernels of a plan run on the FP32
 * pipe (Qmul as one FFMA with a magic constant, exact FADD pre-adds) against the integer form (Qmul_s::mul / Qadd_s::add,
 * include/QuBLAS.h:3146-3235), over EVERY pair of operand raw values the plan's types admit (config 3: all 2^32 pairs,
 * with and without the per-instance clamp elision; config 4: a+b, b-a and the three TFComplexMul products).
 * what: 0 = qb_gemul_plan, 1 = qb_cgemul_plan.  checked = pairs compared (0: the plan uses no FP-pipe quantizer). */
int qb200_selftest_fp(int what, const void *plan, uint64_t *checked, uint64_t *mismatches, const qb_exec *exec);

/* One call, several GPUs, device-resident shards: shard i = `batch` instances whose buffers live on `device` (or in host
 * memory when exec->mem_kind == QB_MEM_HOST).  The shards run concurrently (one host thread each) and are complete on
 * return.  exec->device is ignored; exec->stream must be NULL. */
typedef struct qb_shard {
    int32_t     device;
    int32_t     reserved;
    const void *A, *B;
    void       *C;
    uint64_t    batch;
} qb_shard;
int qb200_gemul_sharded(const qb_gemul_plan *plan, const qb_shard *shards, uint32_t n_shards, const qb_exec *exec);

/* Optional output all-gather over NVLink / NVSwitch (direct peer copies; not on the compute path):
 * shard i = `rows` rows of width[i] bytes (pitch src_pitch[i]) at src[i] on devices[i]; afterwards it sits at byte offset
 * dst_off[i] of every row of dst[j] (pitch dst_pitch) on EVERY device j.  rows == 1: contiguous shards (batches; src_pitch
 * may be NULL).  rows = N: M-row panels of a column-major M x N matrix (width = panel rows * word, dst_pitch = M * word).
 * seconds (optional): device-timed duration, max over the source devices. */
int qb200_allgather(void *const *dst, const void *const *src, const int32_t *devices, uint32_t n, uint64_t rows, uint64_t dst_pitch,
                    const uint64_t *dst_off, const uint64_t *src_pitch, const uint64_t *width, double *seconds);

/* GEMM -> all-gather in ONE kernel per GPU: shard r (shards[r].device, .A = its M-row panel of op(A), .B = all of B; .C and
 * .batch are ignored) computes rows [r * plan->M, (r + 1) * plan->M) of a (n * plan->M) x N column-major C, and the
 * tensor-core epilogue stores every element into c_full[j] on EVERY device j (its own through local stores, the others'
 * through peer stores over NVLink / NVSwitch), tile by tile behind the MMAs of the next tile.  No copy pass afterwards.
 * Tensor-core plans only (lossless int8-range accumulation, plan->M a multiple of 256), n <= 8; complete on return. */
int qb200_gemul_panels_gathered(const qb_gemul_plan *plan, const qb_shard *shards, uint32_t n, void *const *c_full, const qb_exec *exec);

/* Contiguous shard [first, first+count) of `total` units for `rank` of `world` (batch or M-panel). */
int qb200_shard_range(uint64_t total, uint32_t world, uint32_t rank, uint64_t *first, uint64_t *count);

#ifdef __cplusplus
}
#endif
#endif /* QB200_H */
