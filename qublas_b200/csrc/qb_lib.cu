// qb_lib.cu — C ABI of libqb200.so (include/qb200.h): plan compilation, kernel selection, launches.
//
// No CPU fallback lives here: every compute entry point launches CUDA kernels or fails.
#include <cuda_runtime.h>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <memory>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>
#include <algorithm>

#include "../../include/qb200.h"
#include "qb_plan.h"
#include "qb_kernels.cuh"
#include "qb_fast.h"
#include "qb_i8gemm.h"
#include "qb_jit.h"
#include "qb_codegen.h"
#include "qb_fast_cgemv.cuh"

using namespace qb;

static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{0};

static int fail(int code, const std::string &m) { g_err = m; return code; }

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return fail(e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver ? QB_ENODEVICE : QB_ECUDA, \
                        std::string(#expr) + ": " + cudaGetErrorString(e__));                       \
    } while (0)

namespace qb { void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); } }

static int post_launch(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(QB_ECUDA, std::string(what) + " launch: " + cudaGetErrorString(e));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return QB_OK;
}

// ---------------------------------------------------------------------------------------------
// execution context + staging of host buffers
// ---------------------------------------------------------------------------------------------
struct Ctx {
    cudaStream_t stream = nullptr;
    bool host = false;
    uint32_t flags = 0;
    std::vector<void *> to_free;
    int init(const qb_exec *ex)
    {
        static int n_dev = -1;
        if (n_dev <= 0) {
            int n = 0;
            cudaError_t e = cudaGetDeviceCount(&n);
            if (e != cudaSuccess || n == 0) { cudaGetLastError(); return fail(QB_ENODEVICE, "no CUDA device available (there is no CPU fallback)"); }
            n_dev = n;
        }
        if (ex && ex->device >= 0) CUDA_TRY(cudaSetDevice(ex->device));
        {
            // keep stream-ordered scratch (cudaMallocAsync) cached in the pool instead of returning it to the
            // driver at every synchronisation: the default release threshold of 0 makes short calls jitter by 2x
            static std::mutex pool_mu;
            static bool pool_done[64] = { false };
            int dev = 0;
            CUDA_TRY(cudaGetDevice(&dev));
            std::lock_guard<std::mutex> l(pool_mu);
            if (!pool_done[dev & 63]) {
                cudaMemPool_t pool;
                if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                    uint64_t thr = UINT64_MAX;
                    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
                }
                cudaGetLastError();
                pool_done[dev & 63] = true;
            }
        }
        if (ex) {
            stream = (cudaStream_t)ex->stream;
            host = ex->mem_kind == QB_MEM_HOST;
            flags = ex->flags;
        }
        return QB_OK;
    }
    int alloc(void **p, size_t bytes)
    {
        CUDA_TRY(cudaMallocAsync(p, bytes ? bytes : 16, stream));
        to_free.push_back(*p);
        return QB_OK;
    }
    // device view of a caller input
    int in(const void *src, size_t bytes, const void **dev)
    {
        if (!host) { *dev = src; return QB_OK; }
        void *d; int rc = alloc(&d, bytes); if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, stream));
        *dev = d; return QB_OK;
    }
    int out(void *dst, size_t bytes, void **dev)
    {
        if (!host) { *dev = dst; return QB_OK; }
        return alloc(dev, bytes);
    }
    int out_done(void *dst, size_t bytes, void *dev)
    {
        if (host) CUDA_TRY(cudaMemcpyAsync(dst, dev, bytes, cudaMemcpyDeviceToHost, stream));
        return QB_OK;
    }
    int finish()
    {
        for (void *p : to_free) cudaFreeAsync(p, stream);
        to_free.clear();
        if (host) CUDA_TRY(cudaStreamSynchronize(stream));
        return QB_OK;
    }
    ~Ctx() { for (void *p : to_free) cudaFreeAsync(p, stream); }
};

// ---------------------------------------------------------------------------------------------
// Host-buffer calls on large batches: a 3-slot pipeline over internal streams so that the H2D copy
// of chunk i+1, the kernel of chunk i and the D2H copy of chunk i-1 overlap (copy engines + SMs).
// Staging buffers are persistent per device.  With pinned caller memory the copies are truly
// asynchronous; with pageable memory the runtime stages them and the overlap is partial.
// ---------------------------------------------------------------------------------------------
struct HostPipe {
    static constexpr int NS = 6;           // slots created; slots() of them are used (QB200_PIPE_SLOTS, default 3)
    std::mutex mu;
    bool ready = false;
    cudaStream_t st[NS];
    cudaEvent_t done[NS];
    char *buf[NS] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
    size_t cap[NS] = { 0, 0, 0, 0, 0, 0 };
    static int slots()
    {
        static int n = 0;
        if (!n) { const char *e = getenv("QB200_PIPE_SLOTS"); const int v = e ? atoi(e) : 0; n = (v >= 2 && v <= NS) ? v : 3; }
        return n;
    }
    char *shared = nullptr; size_t shared_cap = 0;
    int init()
    {
        if (ready) return QB_OK;
        for (int i = 0; i < NS; i++) {
            CUDA_TRY(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
        }
        ready = true;
        return QB_OK;
    }
    int ensure(char **b, size_t *c, size_t bytes)
    {
        if (*c >= bytes) return QB_OK;
        if (*b) CUDA_TRY(cudaFree(*b));
        *b = nullptr; *c = 0;
        CUDA_TRY(cudaMalloc((void **)b, bytes));
        *c = bytes;
        return QB_OK;
    }
};
static HostPipe g_pipes[64];

struct PipeIn { const char *src; size_t unit_bytes; bool shared; size_t shared_bytes; };

template <class LaunchFn>
static int host_pipeline(int nin, const PipeIn *ins, char *out, size_t out_unit_bytes, uint64_t units, LaunchFn launch)
{
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    HostPipe &hp = g_pipes[dev & 63];
    std::lock_guard<std::mutex> lock(hp.mu);
    int rc = hp.init(); if (rc) return rc;
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    size_t unit_total = out_unit_bytes, shared_total = 0;
    for (int i = 0; i < nin; i++) { if (ins[i].shared) shared_total += al(ins[i].shared_bytes); else unit_total += ins[i].unit_bytes; }
    // Chunk = bytes in + out per pipeline step.  Measured on B200 / PCIe gen5 (tools/e2e_sweep.py): the link saturates with
    // 48..96 MiB chunks as long as a call is cut into >= 8 of them; smaller chunks pay per-copy latency, fewer chunks
    // leave the ends of the pipeline exposed.  QB200_PIPE_MB overrides.
    static long forced_mb = -1;
    if (forced_mb < 0) { const char *e = getenv("QB200_PIPE_MB"); long v = e ? atol(e) : 0; forced_mb = (v > 0 && v <= 1024) ? v : 0; }
    size_t chunk_bytes = (size_t)forced_mb << 20;
    if (!chunk_bytes) {
        const size_t total = (size_t)units * unit_total;
        chunk_bytes = std::min<size_t>((size_t)96 << 20, std::max<size_t>((size_t)12 << 20, total / 8));
    }
    uint64_t cu = chunk_bytes / (unit_total ? unit_total : 1);
    if (cu < 1) cu = 1;
    if (cu > units) cu = units;
    size_t slot_bytes = al(cu * out_unit_bytes);
    for (int i = 0; i < nin; i++) if (!ins[i].shared) slot_bytes += al(cu * ins[i].unit_bytes);
    for (int s = 0; s < HostPipe::slots(); s++) if ((rc = hp.ensure(&hp.buf[s], &hp.cap[s], slot_bytes))) return rc;
    const void *dins[8];
    if (shared_total) {
        if ((rc = hp.ensure(&hp.shared, &hp.shared_cap, shared_total))) return rc;
        size_t off = 0;
        for (int i = 0; i < nin; i++) if (ins[i].shared) {
            CUDA_TRY(cudaMemcpyAsync(hp.shared + off, ins[i].src, ins[i].shared_bytes, cudaMemcpyHostToDevice, hp.st[0]));
            dins[i] = hp.shared + off; off += al(ins[i].shared_bytes);
        }
        CUDA_TRY(cudaStreamSynchronize(hp.st[0]));
    }
    uint64_t first = 0;
    for (uint64_t it = 0; first < units; ++it, first += cu) {
        const int s = (int)(it % HostPipe::slots());
        const uint64_t n = units - first < cu ? units - first : cu;
        cudaStream_t st = hp.st[s];
        size_t off = 0;
        for (int i = 0; i < nin; i++) if (!ins[i].shared) {
            CUDA_TRY(cudaMemcpyAsync(hp.buf[s] + off, ins[i].src + first * ins[i].unit_bytes, n * ins[i].unit_bytes, cudaMemcpyHostToDevice, st));
            dins[i] = hp.buf[s] + off; off += al(cu * ins[i].unit_bytes);
        }
        void *dout = hp.buf[s] + off;
        if ((rc = launch(dins, dout, n, st))) return rc;
        CUDA_TRY(cudaMemcpyAsync(out + first * out_unit_bytes, dout, n * out_unit_bytes, cudaMemcpyDeviceToHost, st));
    }
    for (int s = 0; s < HostPipe::slots(); s++) CUDA_TRY(cudaStreamSynchronize(hp.st[s]));
    return QB_OK;
}

static inline unsigned grid_for(uint64_t n, unsigned block, unsigned cap = 148 * 16)
{
    uint64_t g = (n + block - 1) / block;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}


static void fill_info(qb_plan_info *info, const Analysis &an, qb_qdesc out_t, int out_wb, const char *kernel, int tc, int spec)
{
    memset(info, 0, sizeof(*info));
    info->arith_bits = arith_bits(an);
    info->max_inter_bits = an.max_bits;
    info->lossless = an.lossless;
    info->tensor_core = tc;
    info->specialized = spec;
    info->out_word_bytes = out_wb;
    info->out_type = out_t;
    snprintf(info->kernel, sizeof(info->kernel), "%s", kernel);
}

// ---------------------------------------------------------------------------------------------
// compiled-plan cache: a plan is compiled (types resolved, intervals analysed, kernel selected) once
// per process; later calls with the same plan bytes cost one hash lookup.
// ---------------------------------------------------------------------------------------------
// A JIT-built kernel is a CUmodule in ONE device's primary context, so every cached plan keeps one slot per device.
static constexpr int QB_MAX_DEV = 64;
struct JitSlots {
    std::once_flag once[QB_MAX_DEV];
    std::shared_ptr<JitKernel> k[QB_MAX_DEV];
    static int here() { int d = 0; if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); d = 0; } return d & (QB_MAX_DEV - 1); }
    template <class Build> JitKernel *get(Build build)
    {
        const int d = here();
        std::call_once(once[d], [&] { k[d] = build(); });
        return (k[d] && k[d]->ok) ? k[d].get() : nullptr;
    }
    bool built_here() const { const int d = here(); return k[d] && k[d]->ok; }
};
struct ReduceC { ReduceProg prog; Analysis an; int bits; const FastReduce *fast; JitSlots jit; };
struct GemulC  { GemulProg prog; Analysis an; int bits; const FastGemul *fast; bool tc; JitSlots jit, jit_words; };   // jit_words: word-access kernel for unaligned buffers
struct CGemulC { CGemulProg prog; Analysis an; int bits; const FastCGemul *fast; JitSlots jit; };

template <class Plan, class Entry, class Build>
static std::shared_ptr<Entry> cached(const Plan &p, Build build)
{
    static std::mutex mu;
    static std::unordered_map<std::string, std::shared_ptr<Entry>> map;
    std::string key((const char *)&p, sizeof(Plan));
    {
        std::lock_guard<std::mutex> l(mu);
        auto it = map.find(key);
        if (it != map.end()) return it->second;
    }
    auto e = std::make_shared<Entry>();
    build(*e);
    std::lock_guard<std::mutex> l(mu);
    if (map.size() > 4096) map.clear();
    map[key] = e;
    return e;
}

static std::shared_ptr<ReduceC> get_reduce(const qb_reduce_plan &p)
{
    return cached<qb_reduce_plan, ReduceC>(p, [&](ReduceC &e) {
        compile_reduce(p, &e.prog, e.an);
        e.bits = arith_bits(e.an); e.fast = nullptr;
        if (!e.an.status) { if (e.bits == 32) sanitize32(e.prog.tree); e.fast = getenv("QB200_NO_AOT") ? nullptr : find_fast_reduce(p, e.prog, e.bits); }
    });
}
static std::shared_ptr<GemulC> get_gemul(const qb_gemul_plan &p)
{
    return cached<qb_gemul_plan, GemulC>(p, [&](GemulC &e) {
        compile_gemul(p, &e.prog, e.an);
        e.bits = arith_bits(e.an); e.fast = nullptr; e.tc = false;
        if (!e.an.status) {
            if (e.bits == 32) { sanitize32_stage(e.prog.mul); sanitize32(e.prog.tree); sanitize32_stage(e.prog.out); }
            e.fast = getenv("QB200_NO_AOT") ? nullptr : find_fast_gemul(p, e.prog, e.bits);   // experiments: force the JIT build of the same source
            e.tc = gemul_tensor_core_ok(p, e.prog, e.an) && i8gemm_shape_ok(p.M, p.N, p.K);
        }
    });
}
static std::shared_ptr<CGemulC> get_cgemul(const qb_cgemul_plan &p)
{
    return cached<qb_cgemul_plan, CGemulC>(p, [&](CGemulC &e) {
        compile_cgemul(p, &e.prog, e.an);
        e.bits = arith_bits(e.an); e.fast = nullptr;
        if (!e.an.status) {
            if (e.bits == 32) { sanitize32(e.prog.tf); sanitize32(e.prog.bcm); sanitize32(e.prog.tree_re); sanitize32(e.prog.tree_im); sanitize32_stage(e.prog.out_re); sanitize32_stage(e.prog.out_im); }
            e.fast = getenv("QB200_NO_AOT") ? nullptr : find_fast_cgemul(p, e.prog, e.bits);
        }
    });
}

// ---------------------------------------------------------------------------------------------
// misc entry points
// ---------------------------------------------------------------------------------------------
extern "C" int qb200_abi_version(void) { return QB200_ABI_VERSION; }
extern "C" const char *qb200_last_error(void) { return g_err.c_str(); }
extern "C" int qb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
extern "C" uint64_t qb200_launch_count(void) { return g_launches.load(); }
extern "C" void qb200_launch_count_reset(void) { g_launches.store(0); }

extern "C" int qb200_shard_range(uint64_t total, uint32_t world, uint32_t rank, uint64_t *first, uint64_t *count)
{
    if (!world || rank >= world || !first || !count) return fail(QB_EINVAL, "bad shard arguments");
    uint64_t base = total / world, rem = total % world;
    *first = rank * base + (rank < rem ? rank : rem);
    *count = base + (rank < rem ? 1 : 0);
    return QB_OK;
}

// ---------------------------------------------------------------------------------------------
// exec.device == QB_DEVICE_ALL: one call drives every visible GPU.  The units of a batched call (instances, vectors) are
// independent, so the batch is cut into contiguous shards (qb200_shard_range), one host thread per device runs the
// ordinary single-device path on its shard (own streams, own staging pipeline), and the call returns when all are done.
// There is no inter-GPU traffic on this path.
// ---------------------------------------------------------------------------------------------
template <class Fn>   // fn(int device, uint64_t first, uint64_t count) -> int
static int on_all_devices(uint64_t units, Fn fn)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return fail(QB_ENODEVICE, "no CUDA device available (there is no CPU fallback)"); }
    const uint32_t world = (uint32_t)std::min<uint64_t>((uint64_t)ndev, units ? units : 1);
    int prev = 0;
    cudaGetDevice(&prev);
    std::vector<int> rcs(world, QB_OK);
    std::vector<std::string> errs(world);
    std::vector<std::thread> th;
    for (uint32_t r = 0; r < world; r++) {
        th.emplace_back([&, r] {
            uint64_t first = 0, count = 0;
            qb200_shard_range(units, world, r, &first, &count);
            rcs[r] = count ? fn((int)r, first, count) : QB_OK;
            if (rcs[r]) errs[r] = g_err;                                   // the error text is thread-local
        });
    }
    for (auto &t : th) t.join();
    cudaSetDevice(prev);
    for (uint32_t r = 0; r < world; r++) if (rcs[r]) return fail(rcs[r], "device " + std::to_string(r) + ": " + errs[r]);
    return QB_OK;
}
static bool wants_all_devices(const qb_exec *ex) { return ex && ex->device == QB_DEVICE_ALL; }
static int check_all_devices(const qb_exec *ex)
{
    if (ex->mem_kind != QB_MEM_HOST) return fail(QB_EINVAL, "QB_DEVICE_ALL shards HOST buffers; device-resident shards go through qb200_gemul_sharded");
    if (ex->stream) return fail(QB_EINVAL, "QB_DEVICE_ALL runs on the library's own streams (exec.stream must be NULL)");
    return QB_OK;
}

// ---------------------------------------------------------------------------------------------
// Qreduce
// ---------------------------------------------------------------------------------------------
static unsigned ceil_log2(uint64_t n) { unsigned l = 0; while (((uint64_t)1 << l) < n) l++; return l; }

static const char *reduce_kernel_name(const qb_reduce_plan *p)
{
    if (p->len == 1) return "recast";
    if (p->tree_form == QB_TREE_VECTOR && p->len >= 64) return "k_reduce_warp";
    return "k_reduce_thread";
}

// With a power-of-two length no level ever has an odd leftover, so the variadic form (QuBLAS.h:4936-4949) and the vector
// form (:4966-4980) are the same computation: run it on the vector-form kernels.
static qb_reduce_plan canonical_reduce(const qb_reduce_plan &p)
{
    qb_reduce_plan q = p;
    if (q.tree_form == QB_TREE_VARIADIC && q.len && !(q.len & (q.len - 1))) q.tree_form = QB_TREE_VECTOR;
    return q;
}

extern "C" int qb200_reduce_plan_info(const qb_reduce_plan *plan_in, qb_plan_info *info)
{
    if (!plan_in || !info) return fail(QB_EINVAL, "null argument");
    const qb_reduce_plan canon = canonical_reduce(*plan_in), *plan = &canon;
    auto ce = get_reduce(*plan);
    if (ce->an.status) return fail(ce->an.status, ce->an.err);
    const ReduceProg &rp = ce->prog; const Analysis &an = ce->an; const FastReduce *fr = ce->fast;
    const JitKernel *jk = ce->jit.built_here() ? ce->jit.k[JitSlots::here()].get() : nullptr;
    fill_info(info, an, rp.out.t, rp.out_wb, fr ? fr->name : (jk ? jk->name.c_str() : reduce_kernel_name(plan)), 0, fr != nullptr || jk != nullptr);
    return QB_OK;
}

template <typename I>
static int launch_reduce_warp(const void *x, int in_wb, void *out, int out_wb, uint64_t n_in, uint64_t batch, uint32_t lev0,
                              uint32_t pw, uint32_t rmax, const TreeProg &P, cudaStream_t s)
{
    const uint64_t per_vec = (n_in + ((uint64_t)1 << pw) - 1) >> pw;
    const uint64_t warps = batch * per_vec;
    const unsigned block = 256;
    const uint64_t blocks = (warps + 7) / 8;
    if (blocks > 0x7fffffffull) return fail(QB_EINVAL, "reduction too large for one launch");
    const int vec_ok = ((uintptr_t)x % 16 == 0) && (batch == 1 || (n_in * in_wb) % 16 == 0);
    switch (in_wb) {
    case 1: k_reduce_warp<int8_t, I><<<(unsigned)blocks, block, 0, s>>>((const int8_t *)x, out, out_wb, n_in, batch, lev0, pw, rmax, 0, P); break;
    case 2: k_reduce_warp<int16_t, I><<<(unsigned)blocks, block, 0, s>>>((const int16_t *)x, out, out_wb, n_in, batch, lev0, pw, rmax, vec_ok, P); break;
    case 4: k_reduce_warp<int32_t, I><<<(unsigned)blocks, block, 0, s>>>((const int32_t *)x, out, out_wb, n_in, batch, lev0, pw, rmax, vec_ok, P); break;
    default: k_reduce_warp<int64_t, I><<<(unsigned)blocks, block, 0, s>>>((const int64_t *)x, out, out_wb, n_in, batch, lev0, pw, rmax, vec_ok, P); break;
    }
    return post_launch("k_reduce_warp");
}

template <typename I>
static int run_reduce(const qb_reduce_plan *plan, const ReduceProg &rp, const void *dx, void *dout, void *dpart,
                      uint32_t partial_level, uint64_t batch, Ctx &cx)
{
    const uint64_t len = plan->len;
    cudaStream_t s = cx.stream;
    if (len == 1) {
        k_recast<<<grid_for(batch, 256), 256, 0, s>>>(dx, rp.elem_wb, dout, rp.out_wb, batch);
        int rc = post_launch("k_recast"); if (rc) return rc;
        if (dpart) { k_recast<<<grid_for(batch, 256), 256, 0, s>>>(dx, rp.elem_wb, dpart, 8, batch); return post_launch("k_recast"); }
        return QB_OK;
    }
    const bool warp_path = plan->tree_form == QB_TREE_VECTOR && (len >= 64 || dpart);
    if (!warp_path) {
        k_reduce_thread<I><<<(unsigned)((batch + 127) / 128), 128, 0, s>>>(dx, rp.elem_wb, dout, rp.out_wb, len, batch, rp.tree);
        return post_launch("k_reduce_thread");
    }
    if (dpart && partial_level == 0) {
        k_recast<<<grid_for(batch * len, 256), 256, 0, s>>>(dx, rp.elem_wb, dpart, 8, batch * len);
        int rc = post_launch("k_recast"); if (rc) return rc;
    }
    const void *in = dx; int in_wb = rp.elem_wb; uint64_t n = len; uint32_t lev = 0;
    while (true) {
        const unsigned stop = ceil_log2(n);
        uint32_t pw = stop < 8 ? 8 : (stop > 12 ? 12 : stop);
        bool part_pass = dpart && partial_level > 0 && lev == 0 && partial_level < stop;
        if (part_pass) pw = partial_level;
        const bool final = n <= ((uint64_t)1 << pw);
        if (final) {
            int rc = launch_reduce_warp<I>(in, in_wb, dout, rp.out_wb, n, batch, lev, pw, stop, rp.tree, s);
            if (rc) return rc;
            if (dpart && partial_level > 0 && lev == 0) {            // requested level is the root itself
                k_recast<<<grid_for(batch, 256), 256, 0, s>>>(dout, rp.out_wb, dpart, 8, batch);
                rc = post_launch("k_recast"); if (rc) return rc;
            }
            return QB_OK;
        }
        const uint64_t n_out = (n + ((uint64_t)1 << pw) - 1) >> pw;
        void *scratch; int rc = cx.alloc(&scratch, batch * n_out * sizeof(I)); if (rc) return rc;
        rc = launch_reduce_warp<I>(in, in_wb, scratch, (int)sizeof(I), n, batch, lev, pw, pw, rp.tree, s);
        if (rc) return rc;
        if (part_pass) {
            k_recast<<<grid_for(batch * n_out, 256), 256, 0, s>>>(scratch, (int)sizeof(I), dpart, 8, batch * n_out);
            rc = post_launch("k_recast"); if (rc) return rc;
        }
        in = scratch; in_wb = (int)sizeof(I); n = n_out; lev += pw;
    }
}

extern "C" int qb200_reduce_batched(const qb_reduce_plan *plan_in, const void *x, void *out, void *partials,
                                    uint32_t partial_level, uint64_t batch, const qb_exec *exec)
{
    if (!plan_in || !x || !out) return fail(QB_EINVAL, "null argument");
    if (plan_in->tree_form > 1) return fail(QB_EINVAL, "bad tree_form");
    const qb_reduce_plan canon = canonical_reduce(*plan_in), *plan = &canon;
    auto ce = get_reduce(*plan);
    if (ce->an.status) return fail(ce->an.status, ce->an.err);
    const ReduceProg &rp = ce->prog; const Analysis &an = ce->an;
    if (partials) {
        if (plan_in->tree_form != QB_TREE_VECTOR) return fail(QB_EINVAL, "partials are defined for the vector form only");
        if (partial_level != 0 && partial_level < 8) return fail(QB_EINVAL, "partial_level must be 0 or >= 8");
        if (partial_level > ceil_log2(plan->len)) return fail(QB_EINVAL, "partial_level above the root");
    }
    if (batch == 0) return QB_OK;
    if (wants_all_devices(exec)) {
        int rc = check_all_devices(exec); if (rc) return rc;
        uint64_t per = plan->len; for (uint32_t l = 0; l < partial_level; l++) per = (per + 1) / 2;
        return on_all_devices(batch, [&](int dev, uint64_t first, uint64_t count) {
            qb_exec ex = *exec; ex.device = dev;
            return qb200_reduce_batched(plan_in, (const char *)x + first * plan->len * rp.elem_wb, (char *)out + first * rp.out_wb,
                                        partials ? (char *)partials + first * per * 8 : nullptr, partial_level, count, &ex);
        });
    }
    Ctx cx; int rc = cx.init(exec); if (rc) return rc;
    uint64_t per = plan->len; for (uint32_t l = 0; l < partial_level; l++) per = (per + 1) / 2;
    const void *dx; void *dout, *dpart = nullptr;
    if ((rc = cx.in(x, batch * plan->len * rp.elem_wb, &dx))) return rc;
    if ((rc = cx.out(out, batch * rp.out_wb, &dout))) return rc;
    if (partials && (rc = cx.out(partials, batch * per * 8, &dpart))) return rc;
    const FastReduce *fr = (partials || (cx.flags & QB_F_NO_SPECIALIZE) || ((uintptr_t)dx & 15)) ? nullptr : ce->fast;
    JitKernel *jk = nullptr;
    if (!fr && !partials && !(cx.flags & QB_F_NO_SPECIALIZE) && !((uintptr_t)dx & 15)) {
        jk = ce->jit.get([&] { return jit_reduce(*plan, rp, arith_bits(an)); });
    }
    if (fr) {
        void *scratch;
        if ((rc = cx.alloc(&scratch, batch * 4096 * sizeof(int32_t)))) return rc;
        int e = fr->launch(dx, dout, scratch, batch, cx.stream);
        rc = e ? fail(QB_ECUDA, std::string(fr->name) + " launch failed: " + cudaGetErrorString((cudaError_t)e)) : QB_OK;
    } else if (jk && jk->npass > 0 && jk->variadic) {
        // whole 256-node chunks pass by pass, then one thread per vector for the < 256 trailing nodes of every pass
        uint64_t roots_words = 0;
        for (int i = 0; i < jk->npass; i++) roots_words += batch * jk->pass[i].n_out;
        void *roots;
        if ((rc = cx.alloc(&roots, roots_words * sizeof(int32_t)))) return rc;
        const int32_t *roots_of[4] = { nullptr, nullptr, nullptr, nullptr };
        int e = jit_launch_reduce_var(*jk, dx, roots, batch, cx.stream, roots_of);
        if (e) return fail(QB_ECUDA, "k_reduce_var<jit> launch failed (CUresult " + std::to_string(e) + ")");
        VarRuns R; memset(&R, 0, sizeof(R));
        const int np = jk->npass;
        R.r[R.n++] = { roots_of[np - 1], jk->pass[np - 1].n_out, 0, jk->pass[np - 1].n_out, (uint32_t)(8 * np), 0 };      // every root of the last pass
        for (int i = np - 1; i >= 0; i--) {                                                                             // then the nodes each pass left behind
            const uint64_t n_i = jk->pass[i].n_in, done = jk->pass[i].n_out << 8;
            if (n_i > done) R.r[R.n++] = { i ? roots_of[i - 1] : (const int32_t *)dx, n_i, done, n_i - done, (uint32_t)(8 * i), 0 };
        }
        const uint64_t blocks = (batch + 127) / 128;
        if (blocks > 0x7fffffffull) return fail(QB_EINVAL, "too many vectors for one launch");
        k_reduce_var_finish<int32_t><<<(unsigned)blocks, 128, 0, cx.stream>>>(R, dout, rp.out_wb, batch, rp.tree);
        rc = post_launch("k_reduce_var_finish");
    } else if (jk && jk->npass > 0) {
        void *scratch = nullptr;
        if (jk->npass > 1 && (rc = cx.alloc(&scratch, 2 * batch * jk->pass[0].n_out * sizeof(int32_t)))) return rc;
        int e = jit_launch_reduce_any(*jk, dx, dout, scratch, batch, cx.stream);
        rc = e ? fail(QB_ECUDA, "k_reduce_any<jit> launch failed (CUresult " + std::to_string(e) + ")") : QB_OK;
    } else if (jk) {
        void *scratch;
        if ((rc = cx.alloc(&scratch, batch * 4096 * sizeof(int32_t)))) return rc;
        unsigned logn = 0; while ((1ull << logn) < plan->len) logn++;
        int e = jit_launch_reduce(*jk, logn, dx, dout, scratch, batch, cx.stream);
        rc = e ? fail(QB_ECUDA, "k_reduce_ct<jit> launch failed (CUresult " + std::to_string(e) + ")") : QB_OK;
    } else {
        rc = arith_bits(an) == 32 ? run_reduce<int32_t>(plan, rp, dx, dout, dpart, partial_level, batch, cx)
                                  : run_reduce<int64_t>(plan, rp, dx, dout, dpart, partial_level, batch, cx);
    }
    if (rc) return rc;
    if ((rc = cx.out_done(out, batch * rp.out_wb, dout))) return rc;
    if (partials && (rc = cx.out_done(partials, batch * per * 8, dpart))) return rc;
    return cx.finish();
}

// ---------------------------------------------------------------------------------------------
// Qgemul (real)
// ---------------------------------------------------------------------------------------------
static GemulShape shape_of(uint32_t M, uint32_t N, uint32_t K, int ta, int tb, int a_wb, int b_wb, int c_wb,
                           int64_t sa, int64_t sb, int64_t sc)
{
    GemulShape S; S.M = M; S.N = N; S.K = K; S.trans_a = ta; S.trans_b = tb;
    S.a_wb = a_wb; S.b_wb = b_wb; S.c_wb = c_wb; S.stride_a = sa; S.stride_b = sb; S.stride_c = sc;
    return S;
}

static size_t span_elems(uint64_t batch, int64_t stride, uint64_t per)
{
    return batch == 0 ? 0 : (size_t)((batch - 1) * (uint64_t)stride + per);
}

extern "C" int qb200_gemul_plan_info(const qb_gemul_plan *plan, qb_plan_info *info)
{
    if (!plan || !info) return fail(QB_EINVAL, "null argument");
    auto ce = get_gemul(*plan);
    if (ce->an.status) return fail(ce->an.status, ce->an.err);
    const GemulProg &gp = ce->prog; const Analysis &an = ce->an; const FastGemul *f = ce->fast;
    qb_qdesc ct = plan->c; ct.word_bytes = 0;
    const bool tc = ce->tc;
    const JitKernel *jk = ce->jit.built_here() ? ce->jit.k[JitSlots::here()].get() : nullptr;
    fill_info(info, an, ct, gp.c_wb, tc ? "k_i8gemm (tcgen05 kind::i8)" : (f ? f->name : (jk ? jk->name.c_str() : "k_gemul_thread")), tc,
              f != nullptr || jk != nullptr);
    return QB_OK;
}

// set by qb200_gemul_panels_gathered around its per-shard call: the tensor-core epilogue stores into every device's C
static thread_local const I8FanOut *tl_fanout = nullptr;

extern "C" int qb200_gemul_batched(const qb_gemul_plan *plan, const void *A, const void *B, void *C, uint64_t batch,
                                   const qb_exec *exec)
{
    if (!plan || !A || !B || !C) return fail(QB_EINVAL, "null argument");
    auto ce = get_gemul(*plan);
    if (ce->an.status) return fail(ce->an.status, ce->an.err);
    const GemulProg &gp = ce->prog;
    if (plan->stride_a < 0 || plan->stride_b < 0 || plan->stride_c < (int64_t)((uint64_t)plan->M * plan->N))
        return fail(QB_EINVAL, "bad batch strides");
    if (batch == 0) return QB_OK;
    if (wants_all_devices(exec)) {
        int rc = check_all_devices(exec); if (rc) return rc;
        return on_all_devices(batch, [&](int dev, uint64_t first, uint64_t count) {
            qb_exec ex = *exec; ex.device = dev;
            return qb200_gemul_batched(plan, (const char *)A + first * plan->stride_a * gp.a_wb, (const char *)B + first * plan->stride_b * gp.b_wb,
                                       (char *)C + first * plan->stride_c * gp.c_wb, count, &ex);
        });
    }
    Ctx cx; int rc = cx.init(exec); if (rc) return rc;
    const uint64_t ea = (uint64_t)plan->M * plan->K, eb = (uint64_t)plan->K * plan->N, ec = (uint64_t)plan->M * plan->N;
    const int bits = ce->bits;
    const FastGemul *f = (cx.flags & QB_F_NO_SPECIALIZE) ? nullptr : ce->fast;
    JitKernel *jk = nullptr;
    const bool dense_plan = plan->stride_a == (int64_t)ea && plan->stride_b == (int64_t)eb && plan->stride_c == (int64_t)ec;
    if ((!f || !dense_plan) && !ce->tc && !(cx.flags & QB_F_NO_SPECIALIZE)) {      // run-time specialised kernels carry the plan's strides
        jk = ce->jit.get([&] { return jit_gemul(*plan, gp, bits); });
    }
    // one launch over n instances laid out with the given strides
    auto launch = [&](const void *dA, const void *dB, void *dC, uint64_t n, int64_t sa, int64_t sb, int64_t sc, cudaStream_t st) -> int {
        // the specialised kernels stage operands with cp.async.bulk and store C as int4: 16-byte aligned buffers only
        const bool aligned16 = !(((uintptr_t)dA | (uintptr_t)dB | (uintptr_t)dC) & 15);
        if (f && aligned16 && sa == (int64_t)ea && sb == (int64_t)eb && sc == (int64_t)ec) {
            int e = f->launch(dA, dB, dC, n, st);
            if (e) return fail(QB_ECUDA, std::string(f->name) + " launch failed: " + cudaGetErrorString((cudaError_t)e));
            return QB_OK;
        }
        if (jk && (aligned16 || jk->outputs_per_unit) && sa == plan->stride_a && sb == plan->stride_b && sc == plan->stride_c) {
            int e = jit_launch_gemul(*jk, dA, dB, dC, n, st);      // the thread-per-output template reads and writes word by word
            if (e) return fail(QB_ECUDA, jk->name + " launch failed (CUresult " + std::to_string(e) + ")");
            return QB_OK;
        }
        if (!aligned16 && !(cx.flags & QB_F_NO_SPECIALIZE) && sa == plan->stride_a && sb == plan->stride_b && sc == plan->stride_c) {
            // buffers that are only word-aligned: the specialised thread-per-output kernel instead of the register tiles
            JitKernel *jw = ce->jit_words.get([&] { return jit_gemul_thread(*plan, gp, bits); });
            if (jw) {
                int e = jit_launch_gemul(*jw, dA, dB, dC, n, st);
                if (e) return fail(QB_ECUDA, jw->name + " launch failed (CUresult " + std::to_string(e) + ")");
                return QB_OK;
            }
        }
        GemulShape S = shape_of(plan->M, plan->N, plan->K, plan->trans_a, plan->trans_b, gp.a_wb, gp.b_wb, gp.c_wb, sa, sb, sc);
        const uint64_t blocks = (n * ec + 127) / 128;
        if (blocks > 0x7fffffffull) return fail(QB_EINVAL, "too many outputs for one launch");
        if (bits == 32) k_gemul_thread<int32_t><<<(unsigned)blocks, 128, 0, st>>>(dA, dB, dC, n, S, gp.mul, gp.tree, gp.out);
        else            k_gemul_thread<int64_t><<<(unsigned)blocks, 128, 0, st>>>(dA, dB, dC, n, S, gp.mul, gp.tree, gp.out);
        return post_launch("k_gemul_thread");
    };
    const bool tc = !(cx.flags & QB_F_NO_TENSOR_CORE) && ce->tc;
    if (tl_fanout && (!tc || batch != 1 || cx.host)) return fail(QB_EINVAL, "the fused all-gather needs one device-resident instance on the tensor-core path");
    if (tc) {
        // lossless accumulation: exact int8 x int8 -> int32 dot products on the tensor cores, C's quantizer in the epilogue
        const size_t bytes_a = span_elems(batch, plan->stride_a, ea) * gp.a_wb, bytes_b = span_elems(batch, plan->stride_b, eb) * gp.b_wb,
                     bytes_c = span_elems(batch, plan->stride_c, ec) * gp.c_wb;
        const void *dA, *dB; void *dC; void *Ap, *Bp;
        if ((rc = cx.in(A, bytes_a, &dA)) || (rc = cx.in(B, bytes_b, &dB)) || (rc = cx.out(C, bytes_c, &dC))) return rc;
        uint32_t Mp, Np, Kp;
        i8gemm_padded_dims(plan->M, plan->N, plan->K, &Mp, &Np, &Kp);
        // staging only for operands that have to be packed (int8 K-major operands are read in place)
        bool need_a = false, need_b = false;
        for (uint64_t i = 0; i < batch; i++) {
            need_a |= !i8gemm_zero_copy((const char *)dA + i * plan->stride_a * gp.a_wb, gp.a_wb, plan->trans_a != 0, plan->K);
            need_b |= !i8gemm_zero_copy((const char *)dB + i * plan->stride_b * gp.b_wb, gp.b_wb, plan->trans_b == 0, plan->K);
        }
        Ap = Bp = nullptr;
        if ((need_a && (rc = cx.alloc(&Ap, (size_t)Mp * Kp))) || (need_b && (rc = cx.alloc(&Bp, (size_t)Np * Kp)))) return rc;
        // flags of the operand narrowing fused into the GEMM kernel (int32 words, K contiguous, tile-multiple shapes)
        size_t sync_words = 0;
        for (uint64_t i = 0; i < batch; i++)
            sync_words = std::max(sync_words, i8gemm_sync_words((const char *)dA + i * plan->stride_a * gp.a_wb, (const char *)dB + i * plan->stride_b * gp.b_wb,
                                                                plan->M, plan->N, plan->K, plan->trans_a, plan->trans_b, gp.a_wb, gp.b_wb));
        void *sync = nullptr;
        if (sync_words && (rc = cx.alloc(&sync, sync_words * sizeof(uint32_t)))) return rc;
        for (uint64_t i = 0; i < batch; i++) {
            const char *err = "";
            if (i8gemm_run((const char *)dA + i * plan->stride_a * gp.a_wb, (const char *)dB + i * plan->stride_b * gp.b_wb,
                           (char *)dC + i * plan->stride_c * gp.c_wb, plan->M, plan->N, plan->K, plan->trans_a, plan->trans_b, gp.a_wb, gp.b_wb,
                           gp.c_wb, gp.out, (int8_t *)Ap, (int8_t *)Bp, (uint32_t *)sync, cx.stream, &err, tl_fanout))
                return fail(QB_ECUDA, std::string("k_i8gemm: ") + err);
        }
        if ((rc = cx.out_done(C, bytes_c, dC))) return rc;
        return cx.finish();
    }
    const bool dense = (plan->stride_a == (int64_t)ea || plan->stride_a == 0) && (plan->stride_b == (int64_t)eb || plan->stride_b == 0) &&
                       plan->stride_c == (int64_t)ec;
    if (cx.host && dense && batch >= 64) {
        PipeIn ins[2] = { { (const char *)A, ea * gp.a_wb, plan->stride_a == 0, ea * gp.a_wb }, { (const char *)B, eb * gp.b_wb, plan->stride_b == 0, eb * gp.b_wb } };
        return host_pipeline(2, ins, (char *)C, ec * gp.c_wb, batch, [&](const void *const *d, void *dC, uint64_t n, cudaStream_t st) {
            return launch(d[0], d[1], dC, n, plan->stride_a, plan->stride_b, plan->stride_c, st);
        });
    }
    const size_t bytes_a = span_elems(batch, plan->stride_a, ea) * gp.a_wb, bytes_b = span_elems(batch, plan->stride_b, eb) * gp.b_wb,
                 bytes_c = span_elems(batch, plan->stride_c, ec) * gp.c_wb;
    const void *dA, *dB; void *dC;
    if ((rc = cx.in(A, bytes_a, &dA)) || (rc = cx.in(B, bytes_b, &dB)) || (rc = cx.out(C, bytes_c, &dC))) return rc;
    if ((rc = launch(dA, dB, dC, batch, plan->stride_a, plan->stride_b, plan->stride_c, cx.stream))) return rc;
    if ((rc = cx.out_done(C, bytes_c, dC))) return rc;
    return cx.finish();
}

// ---------------------------------------------------------------------------------------------
// complex Qgemul with TFComplexMul products
// ---------------------------------------------------------------------------------------------
extern "C" int qb200_cgemul_plan_info(const qb_cgemul_plan *plan, qb_plan_info *info)
{
    if (!plan || !info) return fail(QB_EINVAL, "null argument");
    auto ce = get_cgemul(*plan);
    if (ce->an.status) return fail(ce->an.status, ce->an.err);
    const CGemulProg &gp = ce->prog; const Analysis &an = ce->an; const FastCGemul *f = ce->fast;
    qb_qdesc ct = plan->c_re; ct.word_bytes = 0;
    const JitKernel *jk = ce->jit.built_here() ? ce->jit.k[JitSlots::here()].get() : nullptr;
    fill_info(info, an, ct, gp.c_wb, f ? f->name : (jk ? jk->name.c_str() : "k_cgemul_thread"), 0, f != nullptr || jk != nullptr);
    return QB_OK;
}

extern "C" int qb200_cgemul_batched(const qb_cgemul_plan *plan, const void *A, const void *B, void *C, uint64_t batch,
                                    const qb_exec *exec)
{
    if (!plan || !A || !B || !C) return fail(QB_EINVAL, "null argument");
    auto ce = get_cgemul(*plan);
    if (ce->an.status) return fail(ce->an.status, ce->an.err);
    const CGemulProg &gp = ce->prog;
    if (plan->stride_a < 0 || plan->stride_b < 0 || plan->stride_c < (int64_t)((uint64_t)plan->M * plan->N))
        return fail(QB_EINVAL, "bad batch strides");
    if (batch == 0) return QB_OK;
    if (wants_all_devices(exec)) {
        int rc = check_all_devices(exec); if (rc) return rc;
        const size_t wpa = plan->product == QB_CPROD_REAL_A ? 1 : 2, wpb = plan->product == QB_CPROD_REAL_B ? 1 : 2;
        return on_all_devices(batch, [&](int dev, uint64_t first, uint64_t count) {
            qb_exec ex = *exec; ex.device = dev;
            return qb200_cgemul_batched(plan, (const char *)A + first * plan->stride_a * wpa * gp.a_wb, (const char *)B + first * plan->stride_b * wpb * gp.b_wb,
                                        (char *)C + first * plan->stride_c * 2 * gp.c_wb, count, &ex);
        });
    }
    Ctx cx; int rc = cx.init(exec); if (rc) return rc;
    const uint64_t ea = (uint64_t)plan->M * plan->K, eb = (uint64_t)plan->K * plan->N, ec = (uint64_t)plan->M * plan->N;
    const size_t wa = plan->product == QB_CPROD_REAL_A ? 1 : 2, wbn = plan->product == QB_CPROD_REAL_B ? 1 : 2;   // words per element
    const size_t bytes_a = wa * span_elems(batch, plan->stride_a, ea) * gp.a_wb, bytes_b = wbn * span_elems(batch, plan->stride_b, eb) * gp.b_wb,
                 bytes_c = 2 * span_elems(batch, plan->stride_c, ec) * gp.c_wb;
    const void *dA, *dB; void *dC;
    if ((rc = cx.in(A, bytes_a, &dA)) || (rc = cx.in(B, bytes_b, &dB)) || (rc = cx.out(C, bytes_c, &dC))) return rc;
    const int bits = ce->bits;
    const FastCGemul *f = (cx.flags & QB_F_NO_SPECIALIZE) ? nullptr : ce->fast;
    JitKernel *jk = nullptr;
    const bool dense_plan = plan->stride_a == (int64_t)ea && plan->stride_b == (int64_t)eb && plan->stride_c == (int64_t)ec;
    if ((!f || !dense_plan) && !(cx.flags & QB_F_NO_SPECIALIZE)) {                  // run-time specialised kernels carry the plan's strides
        jk = ce->jit.get([&] { return jit_cgemul(*plan, gp, bits); });
    }
    const bool aligned16 = !(((uintptr_t)dA | (uintptr_t)dB | (uintptr_t)dC) & 15);      // vector loads / stores of {re, im} pairs
    if (f && dense_plan && aligned16) {
        rc = f->launch(dA, dB, dC, batch, cx.stream);
        if (rc) return fail(QB_ECUDA, std::string(f->name) + " launch failed: " + cudaGetErrorString((cudaError_t)rc));
    } else if (jk && aligned16) {
        int e = jit_launch_gemul(*jk, dA, dB, dC, batch, cx.stream);
        if (e) return fail(QB_ECUDA, jk->name + " launch failed (CUresult " + std::to_string(e) + ")");
    } else {
        GemulShape S = shape_of(plan->M, plan->N, plan->K, plan->trans_a, plan->trans_b, gp.a_wb, gp.b_wb, gp.c_wb,
                                plan->stride_a, plan->stride_b, plan->stride_c);
        const uint64_t threads = batch * ec;
        const uint64_t blocks = (threads + 127) / 128;
        if (blocks > 0x7fffffffull) return fail(QB_EINVAL, "too many outputs for one launch");
        if (bits == 32) k_cgemul_thread<int32_t><<<(unsigned)blocks, 128, 0, cx.stream>>>(dA, dB, dC, batch, S, gp.kind, gp.tf, gp.bcm, gp.tree_re, gp.tree_im, gp.out_re, gp.out_im);
        else            k_cgemul_thread<int64_t><<<(unsigned)blocks, 128, 0, cx.stream>>>(dA, dB, dC, batch, S, gp.kind, gp.tf, gp.bcm, gp.tree_re, gp.tree_im, gp.out_re, gp.out_im);
        if ((rc = post_launch("k_cgemul_thread"))) return rc;
    }
    if ((rc = cx.out_done(C, bytes_c, dC))) return rc;
    return cx.finish();
}

// ---------------------------------------------------------------------------------------------
// element-wise TFComplexMul, conversion, synthetic fill
// ---------------------------------------------------------------------------------------------
// The element-wise kernels do not depend on the length, so their specialised builds are cached on the plan with len = 0.
struct EwC { JitSlots jit; };
template <class Plan> static std::shared_ptr<EwC> ew_cache(Plan key)
{
    key.len = 0;
    return cached<Plan, EwC>(key, [](EwC &) {});
}
static bool ew_jit_usable(const Ctx &cx, std::initializer_list<const void *> ptrs)
{
    if (cx.flags & QB_F_NO_SPECIALIZE) return false;
    for (const void *p : ptrs) if ((uintptr_t)p & 15) return false;       // vector accesses of {re, im} pairs
    return true;
}

extern "C" int qb200_tfcmul_batched(const qb_tfcmul_plan *plan, const void *x, const void *y, void *out, uint64_t batch,
                                    const qb_exec *exec)
{
    if (!plan || !x || !y || !out) return fail(QB_EINVAL, "null argument");
    TfProg tf; BcmProg bcm; TVal re, im; Analysis an;
    if (compile_tfcmul(*plan, &tf, &bcm, &re, &im, an)) return fail(an.status, an.err);
    const uint64_t n = plan->len * batch;
    if (n == 0) return QB_OK;
    Ctx cx; int rc = cx.init(exec); if (rc) return rc;
    const bool real_x = plan->product == QB_CPROD_REAL_A, real_y = plan->product == QB_CPROD_REAL_B;
    const int xw = real_x ? word_bytes(plan->x_re) : std::max(word_bytes(plan->x_re), word_bytes(plan->x_im));
    const int yw = real_y ? word_bytes(plan->y_re) : std::max(word_bytes(plan->y_re), word_bytes(plan->y_im));
    const int ow = std::max(natural_word(re.t), natural_word(im.t));
    const void *dx, *dy; void *dout;
    if ((rc = cx.in(x, (real_x ? 1 : 2) * n * xw, &dx)) || (rc = cx.in(y, (real_y ? 1 : 2) * n * yw, &dy)) || (rc = cx.out(out, 2 * n * ow, &dout))) return rc;
    const int bits = arith_bits(an);
    if (bits == 32) { sanitize32(tf); sanitize32(bcm); }
    JitKernel *jk = nullptr;
    if (ew_jit_usable(cx, { dx, dy, dout }))
        jk = ew_cache(*plan)->jit.get([&] { return jit_ew((int)plan->product, &tf, &bcm, 0, nullptr, xw, yw, ow, bits); });
    if (jk) {
        int e = jit_launch_ew(*jk, dx, dy, dout, n, cx.stream);
        if (e) return fail(QB_ECUDA, jk->name + " launch failed (CUresult " + std::to_string(e) + ")");
    } else {
        if (bits == 32) k_tfcmul<int32_t><<<grid_for(n, 256), 256, 0, cx.stream>>>(dx, dy, dout, n, xw, yw, ow, (int)plan->product, tf, bcm);
        else            k_tfcmul<int64_t><<<grid_for(n, 256), 256, 0, cx.stream>>>(dx, dy, dout, n, xw, yw, ow, (int)plan->product, tf, bcm);
        if ((rc = post_launch("k_tfcmul"))) return rc;
    }
    if ((rc = cx.out_done(out, 2 * n * ow, dout))) return rc;
    return cx.finish();
}

extern "C" int qb200_convert_batched(const qb_convert_plan *plan, const void *in, void *out, uint64_t batch, const qb_exec *exec)
{
    if (!plan || !in || !out) return fail(QB_EINVAL, "null argument");
    Analysis an;
    if (!check_type(plan->from, an, "from type") || !check_type(plan->to, an, "to type")) return fail(an.status, an.err);
    TVal o;
    QStage q = make_stage(plan->from.frac_bits, type_range(plan->from), width(plan->from), plan->to, &o, an);
    if (an.status) return fail(an.status, an.err);
    if (an.max_bits > 126) return fail(QB_EDOMAIN, "conversion intermediate too wide");
    an.max_reg_bits = std::max(an.max_reg_bits, width(plan->from));      // the input word itself lives in a register
    const uint64_t n = plan->len * batch;
    if (n == 0) return QB_OK;
    Ctx cx; int rc = cx.init(exec); if (rc) return rc;
    const int iw = word_bytes(plan->from), ow = word_bytes(plan->to);
    const void *din; void *dout;
    if ((rc = cx.in(in, n * iw, &din)) || (rc = cx.out(out, n * ow, &dout))) return rc;
    const int bits = arith_bits(an);
    if (bits == 32) sanitize32_stage(q);
    JitKernel *jk = nullptr;
    if (ew_jit_usable(cx, { din, dout })) {
        AddStage st; memset(&st, 0, sizeof(st)); st.q = q;
        jk = ew_cache(*plan)->jit.get([&] { return jit_ew(-1, nullptr, nullptr, 3, &st, iw, iw, ow, bits); });
    }
    if (jk) {
        int e = jit_launch_ew(*jk, din, din, dout, n, cx.stream);
        if (e) return fail(QB_ECUDA, jk->name + " launch failed (CUresult " + std::to_string(e) + ")");
    } else {
        if (bits == 32) k_convert<int32_t><<<grid_for(n, 256), 256, 0, cx.stream>>>(din, dout, n, iw, ow, q);
        else            k_convert<int64_t><<<grid_for(n, 256), 256, 0, cx.stream>>>(din, dout, n, iw, ow, q);
        if ((rc = post_launch("k_convert"))) return rc;
    }
    if ((rc = cx.out_done(out, n * ow, dout))) return rc;
    return cx.finish();
}

extern "C" int qb200_synth_fill(void *dst, uint64_t count, uint32_t wb, int64_t lo, int64_t hi, uint64_t seed,
                                uint64_t stream_id, uint64_t first_index, const qb_exec *exec)
{
    if (!dst) return fail(QB_EINVAL, "null argument");
    if (wb != 1 && wb != 2 && wb != 4 && wb != 8) return fail(QB_EINVAL, "word_bytes must be 1/2/4/8");
    if (hi < lo) return fail(QB_EINVAL, "hi < lo");
    if (count == 0) return QB_OK;
    Ctx cx; int rc = cx.init(exec); if (rc) return rc;
    void *d;
    if ((rc = cx.out(dst, count * wb, &d))) return rc;
    const uint64_t key = mix64(seed * 0xD1342543DE82EF95ull + stream_id);
    const uint64_t span = (uint64_t)(hi - lo) + 1;
    k_synth_fill<<<grid_for(count, 256), 256, 0, cx.stream>>>(d, count, (int)wb, lo, span, key, first_index);
    if ((rc = post_launch("k_synth_fill"))) return rc;
    if ((rc = cx.out_done(dst, count * wb, d))) return rc;
    return cx.finish();
}

// ---------------------------------------------------------------------------------------------
// diagnostic: INT32 issue-rate peak (roofline denominator of the integer-ALU-bound kernels)
// ---------------------------------------------------------------------------------------------
template <int NFMA>   // of 8 chains, NFMA are IMAD chains, the rest IADD3+LOP3 chains
__global__ void __launch_bounds__(256) k_int_peak(int *out, int a, int b, int iters)
{
    int x[8];
#pragma unroll
    for (int c = 0; c < 8; c++) x[c] = threadIdx.x + c * 7 + a;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int c = 0; c < 8; c++) {
                if (c < NFMA) x[c] = x[c] * a + b;                 // IMAD
                else x[c] = ((x[c] + a) ^ b) + it;                 // IADD3 + LOP3 (2 ALU ops)
            }
        }
    }
    int s = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) s ^= x[c];
    if (s == 0x7fffffff) out[0] = s;
}

extern "C" int qb200_measure_int_peak(double fma_share, double *tops, const qb_exec *exec)
{
    if (!tops) return fail(QB_EINVAL, "null argument");
    Ctx cx; int rc = cx.init(exec); if (rc) return rc;
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int *d; if ((rc = cx.alloc((void **)&d, 64))) return rc;
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
    const int nfma = fma_share <= 0.0 ? 0 : (fma_share >= 1.0 ? 8 : (fma_share < 0.4 ? 2 : (fma_share < 0.6 ? 4 : 6)));
    const int iters = 2048, grid = sms * 8;
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        CUDA_TRY(cudaEventRecord(e0, cx.stream));
        switch (nfma) {
        case 0: k_int_peak<0><<<grid, 256, 0, cx.stream>>>(d, 3, 5, iters); break;
        case 2: k_int_peak<2><<<grid, 256, 0, cx.stream>>>(d, 3, 5, iters); break;
        case 4: k_int_peak<4><<<grid, 256, 0, cx.stream>>>(d, 3, 5, iters); break;
        case 6: k_int_peak<6><<<grid, 256, 0, cx.stream>>>(d, 3, 5, iters); break;
        default: k_int_peak<8><<<grid, 256, 0, cx.stream>>>(d, 3, 5, iters); break;
        }
        if ((rc = post_launch("k_int_peak"))) return rc;
        CUDA_TRY(cudaEventRecord(e1, cx.stream));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0; CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double ops_per_thread = (double)iters * 8 * (nfma + 2 * (8 - nfma));
        const double t = ops_per_thread * grid * 256 / (ms * 1e-3) / 1e12;
        if (t > best) best = t;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *tops = best;
    return cx.finish();
}

// ---------------------------------------------------------------------------------------------
// diagnostic: compile (NVRTC only, no device needed) the run-time specialisation of a Qgemul plan
// ---------------------------------------------------------------------------------------------
extern "C" int qb200_jit_check_gemul(const qb_gemul_plan *plan, char *log, uint64_t log_cap)
{
    if (!plan) return fail(QB_EINVAL, "null argument");
    auto ce = get_gemul(*plan);
    if (ce->an.status) return fail(ce->an.status, ce->an.err);
    std::string l;
    int rc = jit_check_gemul(*plan, ce->prog, ce->bits, l);
    if (log && log_cap) { snprintf(log, (size_t)log_cap, "%s", l.c_str()); }
    return rc;
}

// Same for the other run-time specialised templates.  what: 1 = qb_cgemul_plan, 2 = qb_reduce_plan, 3 = qb_tfcmul_plan,
// 4 = qb_ew_plan, 5 = qb_convert_plan.  Returns 0 = cubin produced, 1 = no template applies, -1 = NVRTC failed, < -1 = plan error.
extern "C" int qb200_jit_check(int what, const void *plan, char *log, uint64_t log_cap)
{
    if (!plan) return fail(QB_EINVAL, "null argument");
    int rc = 1; std::string l;
    auto run = [&](auto build) { jit_set_compile_only(true); build(); rc = jit_compile_only_result(l); jit_set_compile_only(false); };
    if (what == 1) {
        auto ce = get_cgemul(*(const qb_cgemul_plan *)plan);
        if (ce->an.status) return fail(ce->an.status, ce->an.err);
        run([&] { jit_cgemul(*(const qb_cgemul_plan *)plan, ce->prog, ce->bits); });
    } else if (what == 2) {
        const qb_reduce_plan canon = canonical_reduce(*(const qb_reduce_plan *)plan);
        auto ce = get_reduce(canon);
        if (ce->an.status) return fail(ce->an.status, ce->an.err);
        run([&] { jit_reduce(canon, ce->prog, arith_bits(ce->an)); });
    } else if (what == 3) {
        const qb_tfcmul_plan &p = *(const qb_tfcmul_plan *)plan;
        TfProg tf; BcmProg bcm; TVal re, im; Analysis an;
        if (compile_tfcmul(p, &tf, &bcm, &re, &im, an)) return fail(an.status, an.err);
        const bool real_x = p.product == QB_CPROD_REAL_A, real_y = p.product == QB_CPROD_REAL_B;
        const int xw = real_x ? word_bytes(p.x_re) : std::max(word_bytes(p.x_re), word_bytes(p.x_im));
        const int yw = real_y ? word_bytes(p.y_re) : std::max(word_bytes(p.y_re), word_bytes(p.y_im));
        const int ow = std::max(natural_word(re.t), natural_word(im.t)), bits = arith_bits(an);
        if (bits == 32) { sanitize32(tf); sanitize32(bcm); }
        run([&] { jit_ew((int)p.product, &tf, &bcm, 0, nullptr, xw, yw, ow, bits); });
    } else if (what == 4) {
        const qb_ew_plan &p = *(const qb_ew_plan *)plan;
        if (p.op > 2) return fail(QB_EINVAL, "bad element-wise op");
        Analysis an;
        if (!check_type(p.a, an, "a type") || !check_type(p.b, an, "b type") || !check_type(p.to, an, "result type")) return fail(an.status, an.err);
        TVal ta = { p.a, type_range(p.a) }, tb = { p.b, type_range(p.b) }, o;
        AddStage st; memset(&st, 0, sizeof(st));
        if (p.op == QB_EW_MUL) st.q = make_mul(ta, tb, p.to, &o, an); else st = make_add(ta, tb, p.to, p.op == QB_EW_SUB, &o, an);
        if (an.status) return fail(an.status, an.err);
        const int bits = arith_bits(an);
        if (bits == 32) sanitize32_stage(st.q);
        run([&] { jit_ew(-1, nullptr, nullptr, (int)p.op, &st, word_bytes(p.a), word_bytes(p.b), word_bytes(p.to), bits); });
    } else if (what == 5) {
        const qb_convert_plan &p = *(const qb_convert_plan *)plan;
        Analysis an;
        if (!check_type(p.from, an, "from type") || !check_type(p.to, an, "to type")) return fail(an.status, an.err);
        TVal o;
        AddStage st; memset(&st, 0, sizeof(st));
        st.q = make_stage(p.from.frac_bits, type_range(p.from), width(p.from), p.to, &o, an);
        if (an.status) return fail(an.status, an.err);
        an.max_reg_bits = std::max(an.max_reg_bits, width(p.from));
        const int bits = arith_bits(an);
        if (bits == 32) sanitize32_stage(st.q);
        run([&] { jit_ew(-1, nullptr, nullptr, 3, &st, word_bytes(p.from), word_bytes(p.from), word_bytes(p.to), bits); });
    } else if (what == 6) {
        const qb_poly_plan &p = *(const qb_poly_plan *)plan;
        PolyProg prog; qb_qdesc ot; Analysis an;
        if (compile_poly(p, &prog, &ot, an)) return fail(an.status, an.err);
        const int bits = arith_bits(an);
        if (bits == 32) sanitize32(prog);
        run([&] { jit_poly(prog, word_bytes(p.x), word_bytes(ot), bits); });
    } else if (what == 7) {
        // Qgemul with a fused polynomial epilogue: plan = { qb_gemul_plan, qb_poly_plan } back to back
        struct Pair { qb_gemul_plan g; qb_poly_plan p; };
        const Pair &pr = *(const Pair *)plan;
        PolyProg prog; qb_qdesc ot; Analysis an;
        if (compile_poly(pr.p, &prog, &ot, an)) return fail(an.status, an.err);
        auto ce = get_gemul(pr.g);
        if (ce->an.status) return fail(ce->an.status, ce->an.err);
        if (arith_bits(an) > ce->bits) { rc = 1; }
        else {
            if (ce->bits == 32) sanitize32(prog);
            run([&] { jit_gemul(pr.g, ce->prog, ce->bits, &prog); });
        }
    } else if (what == 8) {
        // Qgemul with the polynomial fused into the loads of both operands: plan = { qb_gemul_plan, qb_poly_plan } back to back
        struct Pair { qb_gemul_plan g; qb_poly_plan p; };
        const Pair &pr = *(const Pair *)plan;
        PolyProg prog; qb_qdesc ot; Analysis an;
        if (compile_poly(pr.p, &prog, &ot, an)) return fail(an.status, an.err);
        auto ce = get_gemul(pr.g);
        if (ce->an.status) return fail(ce->an.status, ce->an.err);
        if (arith_bits(an) > ce->bits) { rc = 1; }
        else {
            if (ce->bits == 32) sanitize32(prog);
            run([&] { jit_gemul(pr.g, ce->prog, ce->bits, nullptr, &prog, &prog); });
        }
    } else {
        return fail(QB_EINVAL, "unknown plan kind");
    }
    if (log && log_cap) snprintf(log, (size_t)log_cap, "%s", l.c_str());
    return rc;
}

// ---------------------------------------------------------------------------------------------
// double -> Qu, Qu -> double, on-device error statistics
// ---------------------------------------------------------------------------------------------
// Qu_s(double) (QuBLAS.h:2387-2393) on the 53-bit mantissa: raw = ovf(round(m * 2^(e + frac))).
__device__ __forceinline__ int64_t from_double_dev(double v, int frac, const QStage &q)
{
    const uint64_t b = (uint64_t)__double_as_longlong(v);
    const int ef = (int)((b >> 52) & 0x7FF);
    if (ef == 0x7FF || (b << 1) == 0) return 0;                         // NaN / inf / zero load as 0 (QuBLAS.h:667-671)
    uint64_t m = b & 0xFFFFFFFFFFFFFull;
    int e;                                                              // value = m * 2^e
    if (ef == 0) e = -1074; else { m |= 0x10000000000000ull; e = ef - 1075; }
    int64_t mi = (b >> 63) ? -(int64_t)m : (int64_t)m;
    const int sh = -(e + frac);                                         // raw = mi * 2^-sh
    int64_t r;
    if (sh > 0) {
        if (sh >= 62) {                                                  // |mi| < 2^53 << 2^(sh-1): everything but floor of a negative gives 0
            r = (q.qu == TRN_TCPL && mi < 0) ? -1 : 0;
        } else r = q_round<int64_t>(mi, sh, q.qu);
    } else if (-sh <= 10) {                                              // |mi| < 2^53: the shifted value still fits 64 bits
        r = (int64_t)((uint64_t)mi << (-sh));
    } else {
        // beyond 64 bits: decide the overflow on the exact value (a 64-bit-wide target holds e.g. -2^63 exactly, and
        // SAT::ZERO must see "out of range", not a pre-clamped bound)
        bool over = true;
        int64_t low = 0;                                                 // low 64 bits of the exact value (what WRP::TCPL keeps)
        if (-sh < 64) {
            const __int128 w = (__int128)mi << (-sh);
            low = (int64_t)(uint64_t)w;
            over = w > (__int128)q.hi || w < (__int128)q.lo;
        }
        if (over) {
            switch (q.of) {
            case SAT_TCPL: case SAT_SMGN: return mi > 0 ? q.hi : q.lo;
            case SAT_ZERO: return 0;
            default: { QStage o = q; o.need_lo = 1; o.need_hi = 1; return q_ovf<int64_t, false>(low, o); }
            }
        }
        r = low;
    }
    QStage o = q; o.need_lo = 1; o.need_hi = 1;
    return q_ovf<int64_t, false>(r, o);
}

__global__ void k_quantize_f64(const double *__restrict__ in, void *__restrict__ out, uint64_t n, int wb, int frac, const __grid_constant__ QStage q)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e0 < n; e0 += 4 * stride) {
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { const uint64_t e = e0 + u * stride; v[u] = e < n ? in[e] : 0.0; }
#pragma unroll
        for (int u = 0; u < 4; u++) { const uint64_t e = e0 + u * stride; if (e < n) store_word<int64_t>(out, e, wb, from_double_dev(v[u], frac, q)); }
    }
}
// 2^-frac as a double (exact for |frac| <= 1022): multiplying by it is the exact scaling of toDouble (QuBLAS.h:2413-2416)
__device__ __forceinline__ double pow2_neg(int frac) { return __longlong_as_double((long long)(1023 - frac) << 52); }

__global__ void k_dequantize_f64(const void *__restrict__ in, double *__restrict__ out, uint64_t n, int wb, int frac)
{
    const double sc = pow2_neg(frac);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e0 < n; e0 += 4 * stride) {
        int64_t v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { const uint64_t e = e0 + u * stride; v[u] = e < n ? load_word<int64_t>(in, e, wb) : 0; }
#pragma unroll
        for (int u = 0; u < 4; u++) { const uint64_t e = e0 + u * stride; if (e < n) out[e] = (wb <= 4 ? (double)(int)v[u] : (double)v[u]) * sc; }   // int32 -> double is one instruction
    }
}
// fixed grid, fixed per-thread order, fixed tree: the sums are reproducible run to run
__global__ void __launch_bounds__(256) k_err_stats(const void *__restrict__ w, const double *__restrict__ ref, uint64_t n, int wb, int frac, double *__restrict__ part)
{
    __shared__ double s0[256], s1[256], s2[256];
    // four independent accumulation chains per thread (loads of four strided elements in flight), combined in a fixed order
    double se4[4] = { 0, 0, 0, 0 }, sr4[4] = { 0, 0, 0, 0 }, mx = 0;
    const double sc = pow2_neg(frac);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e0 < n; e0 += 4 * stride) {
        int64_t wv[4]; double rv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { const uint64_t e = e0 + u * stride; const bool ok = e < n; wv[u] = ok ? load_word<int64_t>(w, e, wb) : 0; rv[u] = ok ? ref[e] : 0.0; }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (e0 + u * stride < n) {
                const double d = (wb <= 4 ? (double)(int)wv[u] : (double)wv[u]) * sc - rv[u];
                se4[u] += d * d; sr4[u] += rv[u] * rv[u]; mx = fmax(mx, fabs(d));
            }
        }
    }
    const double se = (se4[0] + se4[1]) + (se4[2] + se4[3]), sr = (sr4[0] + sr4[1]) + (sr4[2] + sr4[3]);
    s0[threadIdx.x] = se; s1[threadIdx.x] = sr; s2[threadIdx.x] = mx;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { s0[threadIdx.x] += s0[threadIdx.x + o]; s1[threadIdx.x] += s1[threadIdx.x + o]; s2[threadIdx.x] = fmax(s2[threadIdx.x], s2[threadIdx.x + o]); }
        __syncthreads();
    }
    if (threadIdx.x == 0) { part[3 * blockIdx.x] = s0[0]; part[3 * blockIdx.x + 1] = s1[0]; part[3 * blockIdx.x + 2] = s2[0]; }
}
// one block: thread t adds the partials t, t + 256, ... in order, then the same fixed tree as above
__global__ void __launch_bounds__(256) k_err_final(const double *__restrict__ part, int nblocks, double *__restrict__ out)
{
    __shared__ double s0[256], s1[256], s2[256];
    double se = 0, sr = 0, mx = 0;
    for (int i = threadIdx.x; i < nblocks; i += 256) { se += part[3 * i]; sr += part[3 * i + 1]; mx = fmax(mx, part[3 * i + 2]); }
    s0[threadIdx.x] = se; s1[threadIdx.x] = sr; s2[threadIdx.x] = mx;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { s0[threadIdx.x] += s0[threadIdx.x + o]; s1[threadIdx.x] += s1[threadIdx.x + o]; s2[threadIdx.x] = fmax(s2[threadIdx.x], s2[threadIdx.x + o]); }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = s0[0]; out[1] = s1[0]; out[2] = s2[0]; }
}

static int stage_for_type(const qb_qdesc &t, QStage *q, const char *what)
{
    Analysis an;
    if (!check_type(t, an, what)) return fail(an.status, an.err);
    TVal o;
    Ival full = { -((i128)1 << 62), ((i128)1 << 62) };
    *q = make_stage(t.frac_bits, full, 64, t, &o, an);                 // bounds / wrap fields of the target type; shift unused
    return QB_OK;
}

extern "C" int qb200_quantize_f64(const qb_qdesc *to, const double *in, void *out, uint64_t n, const qb_exec *exec)
{
    if (!to || !in || !out) return fail(QB_EINVAL, "null argument");
    QStage q; int rc = stage_for_type(*to, &q, "target type"); if (rc) return rc;
    if (n == 0) return QB_OK;
    Ctx cx; if ((rc = cx.init(exec))) return rc;
    const int wb = word_bytes(*to);
    const void *din; void *dout;
    if ((rc = cx.in(in, n * 8, &din)) || (rc = cx.out(out, n * wb, &dout))) return rc;
    k_quantize_f64<<<grid_for(n, 256), 256, 0, cx.stream>>>((const double *)din, dout, n, wb, to->frac_bits, q);
    if ((rc = post_launch("k_quantize_f64"))) return rc;
    if ((rc = cx.out_done(out, n * wb, dout))) return rc;
    return cx.finish();
}

extern "C" int qb200_dequantize_f64(const qb_qdesc *from, const void *in, double *out, uint64_t n, const qb_exec *exec)
{
    if (!from || !in || !out) return fail(QB_EINVAL, "null argument");
    Analysis an; if (!check_type(*from, an, "source type")) return fail(an.status, an.err);
    if (n == 0) return QB_OK;
    Ctx cx; int rc = cx.init(exec); if (rc) return rc;
    const int wb = word_bytes(*from);
    const void *din; void *dout;
    if ((rc = cx.in(in, n * wb, &din)) || (rc = cx.out(out, n * 8, &dout))) return rc;
    k_dequantize_f64<<<grid_for(n, 256), 256, 0, cx.stream>>>(din, (double *)dout, n, wb, from->frac_bits);
    if ((rc = post_launch("k_dequantize_f64"))) return rc;
    if ((rc = cx.out_done(out, n * 8, dout))) return rc;
    return cx.finish();
}

extern "C" int qb200_error_stats(const qb_qdesc *type, const void *words, const double *ref, uint64_t n, qb_err_stats *stats, const qb_exec *exec)
{
    if (!type || !words || !ref || !stats) return fail(QB_EINVAL, "null argument");
    Analysis an; if (!check_type(*type, an, "type")) return fail(an.status, an.err);
    memset(stats, 0, sizeof(*stats));
    stats->n = n;
    if (n == 0) return QB_OK;
    Ctx cx; int rc = cx.init(exec); if (rc) return rc;
    const int wb = word_bytes(*type), nblocks = 1184;
    const void *dw, *dr; void *part, *res;
    if ((rc = cx.in(words, n * wb, &dw)) || (rc = cx.in(ref, n * 8, &dr)) || (rc = cx.alloc(&part, 3 * nblocks * 8)) || (rc = cx.alloc(&res, 3 * 8))) return rc;
    k_err_stats<<<nblocks, 256, 0, cx.stream>>>(dw, (const double *)dr, n, wb, type->frac_bits, (double *)part);
    if ((rc = post_launch("k_err_stats"))) return rc;
    k_err_final<<<1, 256, 0, cx.stream>>>((const double *)part, nblocks, (double *)res);
    if ((rc = post_launch("k_err_final"))) return rc;
    double h[3];
    CUDA_TRY(cudaMemcpyAsync(h, res, sizeof(h), cudaMemcpyDeviceToHost, cx.stream));
    CUDA_TRY(cudaStreamSynchronize(cx.stream));
    stats->sum_sq_err = h[0]; stats->sum_sq_ref = h[1]; stats->max_abs_err = h[2];
    return cx.finish();
}

// ---------------------------------------------------------------------------------------------
// element-wise Qmul / Qadd / Qsub on tensors
// ---------------------------------------------------------------------------------------------
extern "C" int qb200_elementwise_batched(const qb_ew_plan *plan, const void *a, const void *b, void *out, uint64_t batch, const qb_exec *exec)
{
    if (!plan || !a || !b || !out) return fail(QB_EINVAL, "null argument");
    if (plan->op > 2) return fail(QB_EINVAL, "bad element-wise op");
    Analysis an;
    if (!check_type(plan->a, an, "a type") || !check_type(plan->b, an, "b type") || !check_type(plan->to, an, "result type")) return fail(an.status, an.err);
    TVal ta = { plan->a, type_range(plan->a) }, tb = { plan->b, type_range(plan->b) }, o;
    AddStage st; memset(&st, 0, sizeof(st));
    if (plan->op == QB_EW_MUL) st.q = make_mul(ta, tb, plan->to, &o, an);
    else st = make_add(ta, tb, plan->to, plan->op == QB_EW_SUB, &o, an);
    if (an.status) return fail(an.status, an.err);
    if (an.ref_width > 128 || an.max_bits > 126) return fail(QB_EDOMAIN, "an intermediate would be wider than 128 bits");
    const uint64_t n = plan->len * batch;
    if (n == 0) return QB_OK;
    Ctx cx; int rc = cx.init(exec); if (rc) return rc;
    const int aw = word_bytes(plan->a), bw = word_bytes(plan->b), ow = word_bytes(plan->to);
    const void *da, *db; void *dout;
    if ((rc = cx.in(a, n * aw, &da)) || (rc = cx.in(b, n * bw, &db)) || (rc = cx.out(out, n * ow, &dout))) return rc;
    const int bits = arith_bits(an);
    if (bits == 32) sanitize32_stage(st.q);
    JitKernel *jk = nullptr;
    if (ew_jit_usable(cx, { da, db, dout }))
        jk = ew_cache(*plan)->jit.get([&] { return jit_ew(-1, nullptr, nullptr, (int)plan->op, &st, aw, bw, ow, bits); });
    if (jk) {
        int e = jit_launch_ew(*jk, da, db, dout, n, cx.stream);
        if (e) return fail(QB_ECUDA, jk->name + " launch failed (CUresult " + std::to_string(e) + ")");
    } else {
        if (bits == 32) k_elementwise<int32_t><<<grid_for(n, 256), 256, 0, cx.stream>>>(da, db, dout, n, aw, bw, ow, (int)plan->op, st);
        else            k_elementwise<int64_t><<<grid_for(n, 256), 256, 0, cx.stream>>>(da, db, dout, n, aw, bw, ow, (int)plan->op, st);
        if ((rc = post_launch("k_elementwise"))) return rc;
    }
    if ((rc = cx.out_done(out, n * ow, dout))) return rc;
    return cx.finish();
}

// ---------------------------------------------------------------------------------------------
// ANUS::Qpoly / Qapprox over tensors, as an epilogue of Qgemul, and Qtable lookups
// ---------------------------------------------------------------------------------------------
static int poly_setup(const qb_poly_plan *plan, PolyProg *prog, qb_qdesc *out_t, Analysis &an)
{
    if (compile_poly(*plan, prog, out_t, an)) return fail(an.status, an.err);
    return QB_OK;
}

extern "C" int qb200_poly_plan_info(const qb_poly_plan *plan, qb_plan_info *info)
{
    if (!plan || !info) return fail(QB_EINVAL, "null argument");
    PolyProg prog; qb_qdesc ot; Analysis an;
    int rc = poly_setup(plan, &prog, &ot, an); if (rc) return rc;
    const int ow = word_bytes(ot);
    ot.word_bytes = 0;
    fill_info(info, an, ot, ow, "k_poly", 0, 0);
    return QB_OK;
}

static int launch_poly(const PolyProg &prog_in, int bits, const void *dx, void *dout, uint64_t n, int xw, int ow, cudaStream_t st)
{
    if (n == 0) return QB_OK;
    if (bits == 32) {
        PolyProg prog = prog_in;
        sanitize32(prog);
        k_poly<int32_t><<<grid_for(n, 256), 256, 0, st>>>(dx, dout, n, xw, ow, prog);
    } else {
        k_poly<int64_t><<<grid_for(n, 256), 256, 0, st>>>(dx, dout, n, xw, ow, prog_in);
    }
    return post_launch("k_poly");
}

// specialised kernel (NVRTC, cached per plan and device) when the buffers allow vector accesses, else the runtime-descriptor one
static int run_poly(const qb_poly_plan &plan, const PolyProg &prog, int bits, const Ctx &cx, const void *dx, void *dout, uint64_t n, int xw, int ow)
{
    JitKernel *jk = nullptr;
    if (ew_jit_usable(cx, { dx, dout })) {
        PolyProg p32 = prog;
        if (bits == 32) sanitize32(p32);
        jk = ew_cache(plan)->jit.get([&] { return jit_poly(p32, xw, ow, bits); });
    }
    if (jk) {
        int e = jit_launch_ew(*jk, dx, dx, dout, n, cx.stream);
        if (e) return fail(QB_ECUDA, jk->name + " launch failed (CUresult " + std::to_string(e) + ")");
        return QB_OK;
    }
    return launch_poly(prog, bits, dx, dout, n, xw, ow, cx.stream);
}

extern "C" int qb200_poly_batched(const qb_poly_plan *plan, const void *x, void *out, uint64_t batch, const qb_exec *exec)
{
    if (!plan || !x || !out) return fail(QB_EINVAL, "null argument");
    PolyProg prog; qb_qdesc ot; Analysis an;
    int rc = poly_setup(plan, &prog, &ot, an); if (rc) return rc;
    const uint64_t n = plan->len * batch;
    if (n == 0) return QB_OK;
    Ctx cx; if ((rc = cx.init(exec))) return rc;
    const int xw = word_bytes(plan->x), ow = word_bytes(ot);
    const void *dx; void *dout;
    if ((rc = cx.in(x, n * xw, &dx)) || (rc = cx.out(out, n * ow, &dout))) return rc;
    if ((rc = run_poly(*plan, prog, arith_bits(an), cx, dx, dout, n, xw, ow))) return rc;
    if ((rc = cx.out_done(out, n * ow, dout))) return rc;
    return cx.finish();
}

// Qgemul followed by the polynomial on every output element, before the result leaves the GPU: the GEMM writes words of
// poly->x (= gemul->c) into C's buffer and k_poly rewrites them in place as words of the polynomial's result type.
extern "C" int qb200_gemul_poly_batched(const qb_gemul_plan *gemul, const qb_poly_plan *poly, const void *A, const void *B, void *C, uint64_t batch,
                                        const qb_exec *exec)
{
    if (!gemul || !poly || !A || !B || !C) return fail(QB_EINVAL, "null argument");
    if (memcmp(&gemul->c, &poly->x, sizeof(qb_qdesc)) != 0) return fail(QB_EINVAL, "the polynomial's x type must be the type Qgemul converts into (gemul->c)");
    if (exec && exec->device == QB_DEVICE_ALL) return fail(QB_EINVAL, "QB_DEVICE_ALL is not supported by the fused call: shard the batch and pass one device per call");
    PolyProg prog; qb_qdesc ot; Analysis an;
    int rc = poly_setup(poly, &prog, &ot, an); if (rc) return rc;
    const int xw = word_bytes(poly->x), ow = word_bytes(ot);
    if (xw != ow) return fail(QB_EINVAL, "the polynomial's result words must be as wide as Qgemul's result words (in-place epilogue)");
    if (gemul->stride_c < (int64_t)((uint64_t)gemul->M * gemul->N)) return fail(QB_EINVAL, "bad batch strides");
    if (batch == 0) return QB_OK;
    Ctx cx; if ((rc = cx.init(exec))) return rc;
    const uint64_t ec = (uint64_t)gemul->M * gemul->N;
    const size_t bytes_c = span_elems(batch, gemul->stride_c, ec) * xw;
    void *dC;
    if ((rc = cx.out(C, bytes_c, &dC))) return rc;
    if (cx.host) {
        // operands from host memory: the GEMM call stages them itself and leaves its result in the device buffer
        const uint64_t ea = (uint64_t)gemul->M * gemul->K, eb = (uint64_t)gemul->K * gemul->N;
        const int aw = word_bytes(gemul->a), bw = word_bytes(gemul->b);
        const void *dA, *dB;
        if ((rc = cx.in(A, span_elems(batch, gemul->stride_a, ea) * aw, &dA)) || (rc = cx.in(B, span_elems(batch, gemul->stride_b, eb) * bw, &dB))) return rc;
        A = dA; B = dB;
    }
    // Fused: the run-time specialised Qgemul kernels apply the polynomial in registers, right after the converting assignment
    // (qb_fast_gemul.cuh / qb_fast_gemul_thread.cuh: Cfg::EPI).  Needs a plan without the tensor-core path, the polynomial in
    // the kernel's register class and 16-byte aligned buffers; everything else runs the two kernels back to back.
    {
        auto ce = get_gemul(*gemul);
        if (ce->an.status) return fail(ce->an.status, ce->an.err);
        const bool tc = !(cx.flags & QB_F_NO_TENSOR_CORE) && ce->tc;
        const bool aligned16 = !(((uintptr_t)A | (uintptr_t)B | (uintptr_t)dC) & 15);
        if (!tc && !(cx.flags & QB_F_NO_SPECIALIZE) && aligned16 && arith_bits(an) <= ce->bits) {
            struct Key { qb_gemul_plan g; qb_poly_plan p; } key;
            memset(&key, 0, sizeof(key)); key.g = *gemul; key.p = *poly; key.p.len = 0;
            PolyProg pe = prog;
            if (ce->bits == 32) sanitize32(pe);
            JitKernel *jk = cached<Key, EwC>(key, [](EwC &) {})->jit.get([&] { return jit_gemul(*gemul, ce->prog, ce->bits, &pe); });
            if (jk) {
                int e = jit_launch_gemul(*jk, A, B, dC, batch, cx.stream);
                if (e) return fail(QB_ECUDA, jk->name + " launch failed (CUresult " + std::to_string(e) + ")");
                if ((rc = cx.out_done(C, bytes_c, dC))) return rc;
                return cx.finish();
            }
        }
    }
    int dev = 0; cudaGetDevice(&dev);
    qb_exec ex = { QB_MEM_DEVICE, dev, (void *)cx.stream, cx.flags, 0 };
    if ((rc = qb200_gemul_batched(gemul, A, B, dC, batch, &ex))) return rc;
    // strided C: the words between instances are not results; run instance by instance only if there is padding
    if (gemul->stride_c == (int64_t)ec) {
        if ((rc = launch_poly(prog, arith_bits(an), dC, dC, batch * ec, xw, ow, cx.stream))) return rc;
    } else {
        for (uint64_t i = 0; i < batch; i++) {
            char *ci = (char *)dC + i * (uint64_t)gemul->stride_c * xw;
            if ((rc = launch_poly(prog, arith_bits(an), ci, ci, ec, xw, ow, cx.stream))) return rc;
        }
    }
    if ((rc = cx.out_done(C, bytes_c, dC))) return rc;
    return cx.finish();
}

// The polynomial applied to an OPERAND of Qgemul before the products (Qgemul(C, Qapprox(A), B): the reference evaluates the
// tensor expression element by element, QuBLAS.h:4829-4897 into :3146-3171).  poly_a / poly_b may be NULL; a polynomial's
// result type must be the operand type of the plan (gemul->a / gemul->b), the caller's buffer holds raw words of the
// polynomial's x type in words of the same size.  Fused: the run-time specialised Qgemul kernels apply it as the operand is
// staged (register tiles: once per element, in the conversion pass) or loaded (thread per output); otherwise the polynomial
// runs as its own launch into a scratch copy of the operand.
extern "C" int qb200_gemul_prologue_batched(const qb_gemul_plan *gemul, const qb_poly_plan *poly_a, const qb_poly_plan *poly_b, const void *A, const void *B,
                                            void *C, uint64_t batch, const qb_exec *exec)
{
    if (!gemul || !A || !B || !C) return fail(QB_EINVAL, "null argument");
    if (!poly_a && !poly_b) return qb200_gemul_batched(gemul, A, B, C, batch, exec);
    if (exec && exec->device == QB_DEVICE_ALL) return fail(QB_EINVAL, "QB_DEVICE_ALL is not supported by the fused call: shard the batch and pass one device per call");
    auto ce = get_gemul(*gemul);
    if (ce->an.status) return fail(ce->an.status, ce->an.err);
    const GemulProg &gp = ce->prog;
    PolyProg prog[2]; Analysis an[2]; int pbits = 32;
    const qb_poly_plan *pp[2] = { poly_a, poly_b };
    const qb_qdesc *want[2] = { &gemul->a, &gemul->b };
    const int wb[2] = { gp.a_wb, gp.b_wb };
    int rc;
    for (int i = 0; i < 2; i++) if (pp[i]) {
        qb_qdesc ot;
        if ((rc = poly_setup(pp[i], &prog[i], &ot, an[i]))) return rc;
        if (ot.int_bits != want[i]->int_bits || ot.frac_bits != want[i]->frac_bits || ot.is_signed != want[i]->is_signed)
            return fail(QB_EINVAL, "the polynomial's result type must be the operand type of the Qgemul plan");
        if (word_bytes(pp[i]->x) != wb[i] || word_bytes(ot) != wb[i]) return fail(QB_EINVAL, "the polynomial's x words and the operand words must have the same size");
        pbits = std::max(pbits, arith_bits(an[i]));
    }
    if (gemul->stride_a < 0 || gemul->stride_b < 0 || gemul->stride_c < (int64_t)((uint64_t)gemul->M * gemul->N)) return fail(QB_EINVAL, "bad batch strides");
    if (batch == 0) return QB_OK;
    Ctx cx; if ((rc = cx.init(exec))) return rc;
    const uint64_t ea = (uint64_t)gemul->M * gemul->K, eb = (uint64_t)gemul->K * gemul->N, ec = (uint64_t)gemul->M * gemul->N;
    const uint64_t na = span_elems(batch, gemul->stride_a, ea), nb = span_elems(batch, gemul->stride_b, eb);
    const size_t bytes_c = span_elems(batch, gemul->stride_c, ec) * gp.c_wb;
    const void *dA, *dB; void *dC;
    if ((rc = cx.in(A, na * gp.a_wb, &dA)) || (rc = cx.in(B, nb * gp.b_wb, &dB)) || (rc = cx.out(C, bytes_c, &dC))) return rc;
    const bool tc = !(cx.flags & QB_F_NO_TENSOR_CORE) && ce->tc;
    const bool aligned16 = !(((uintptr_t)dA | (uintptr_t)dB | (uintptr_t)dC) & 15);
    if (!tc && !(cx.flags & QB_F_NO_SPECIALIZE) && aligned16 && pbits <= ce->bits) {
        struct Key { qb_gemul_plan g; qb_poly_plan a, b; uint32_t has_a, has_b; } key;
        memset(&key, 0, sizeof(key)); key.g = *gemul; key.has_a = poly_a != nullptr; key.has_b = poly_b != nullptr;
        if (poly_a) { key.a = *poly_a; key.a.len = 0; }
        if (poly_b) { key.b = *poly_b; key.b.len = 0; }
        PolyProg pe[2] = { prog[0], prog[1] };
        if (ce->bits == 32) { if (poly_a) sanitize32(pe[0]); if (poly_b) sanitize32(pe[1]); }
        JitKernel *jk = cached<Key, EwC>(key, [](EwC &) {})->jit.get([&] { return jit_gemul(*gemul, gp, ce->bits, nullptr, poly_a ? &pe[0] : nullptr, poly_b ? &pe[1] : nullptr); });
        if (jk) {
            int e = jit_launch_gemul(*jk, dA, dB, dC, batch, cx.stream);
            if (e) return fail(QB_ECUDA, jk->name + " launch failed (CUresult " + std::to_string(e) + ")");
            if ((rc = cx.out_done(C, bytes_c, dC))) return rc;
            return cx.finish();
        }
    }
    // not fused: the polynomial into a scratch copy of the operand, then the plain call on device memory
    const void *src[2] = { dA, dB }; const uint64_t n[2] = { na, nb };
    for (int i = 0; i < 2; i++) if (pp[i]) {
        void *tmp;
        if ((rc = cx.alloc(&tmp, n[i] * wb[i])) || (rc = launch_poly(prog[i], arith_bits(an[i]), src[i], tmp, n[i], wb[i], wb[i], cx.stream))) return rc;
        src[i] = tmp;
    }
    int dev = 0; cudaGetDevice(&dev);
    qb_exec ex = { QB_MEM_DEVICE, dev, (void *)cx.stream, cx.flags, 0 };
    if ((rc = qb200_gemul_batched(gemul, src[0], src[1], dC, batch, &ex))) return rc;
    if ((rc = cx.out_done(C, bytes_c, dC))) return rc;
    return cx.finish();
}

extern "C" int qb200_table_batched(const qb_qdesc *x, const qb_qdesc *out_type, const void *table, const void *in, void *out, uint64_t n, const qb_exec *exec)
{
    if (!x || !out_type || !table || !in || !out) return fail(QB_EINVAL, "null argument");
    Analysis an;
    if (!check_type(*x, an, "x type") || !check_type(*out_type, an, "result type")) return fail(an.status, an.err);
    if (width(*x) > 20) return fail(QB_EDOMAIN, "lookup tables are limited to inputs of at most 20 bits");
    if (n == 0) return QB_OK;
    const Ival r = type_range(*x);
    const uint64_t entries = (uint64_t)(r.hi - r.lo + 1);
    Ctx cx; int rc = cx.init(exec); if (rc) return rc;
    const int xw = word_bytes(*x), ow = word_bytes(*out_type);
    const void *din, *dtab; void *dout;
    if ((rc = cx.in(in, n * xw, &din)) || (rc = cx.in(table, entries * ow, &dtab)) || (rc = cx.out(out, n * ow, &dout))) return rc;
    k_table<<<grid_for(n, 256), 256, 0, cx.stream>>>(din, dout, dtab, n, xw, ow, (int64_t)r.lo, entries);
    if ((rc = post_launch("k_table"))) return rc;
    if ((rc = cx.out_done(out, n * ow, dout))) return rc;
    return cx.finish();
}

// ---------------------------------------------------------------------------------------------
// diagnostic / parity evidence: exhaustive sweep of the quantizers that run on the FP32 pipe
// ---------------------------------------------------------------------------------------------
// The specialised kernels replace some integer quantizers by floating-point instructions (qb_device.cuh:fp_mul_apply,
// qb_fast_cgemv.cuh:fadd_stage).  These kernels run the SAME device functions over every operand pair the plan's types
// admit and compare with the integer form (mul_apply / add_apply, the restatement of QuBLAS.h:3146-3235).
// mode 0: clamps as the plan has them; mode 1: clamp-free variant, counted only where the product lies inside [p_lo, p_hi]
// (the per-instance clamp elision of qb_fast_gemul.cuh); mode 2: clamp-free variant on operands whose minima were replaced by
// the substitutes xs / ys (saturation by operand substitution, qb_device.cuh:fp_subst), every pair counted.
template <int S>
__global__ void __launch_bounds__(256) k_selftest_fpmul(long long xlo, unsigned long long nx, long long ylo, unsigned long long ny,
                                                        const __grid_constant__ QStage q, int mode, long long p_lo, long long p_hi,
                                                        float xs, float ys, unsigned long long *counts)
{
    unsigned long long bad = 0, seen = 0;
    QStage qq = q;
    if (mode != 0) { qq.need_lo = 0; qq.need_hi = 0; }
    for (unsigned long long iy = blockIdx.x; iy < ny; iy += gridDim.x) {
        const int y = (int)(ylo + (long long)iy);
        const float yf = mode == 2 ? fmaxf(__int2float_rn(y), ys) : __int2float_rn(y);
        for (unsigned long long ix = threadIdx.x; ix < nx; ix += blockDim.x) {
            const int x = (int)(xlo + (long long)ix);
            if (mode == 1) { const long long p = (long long)x * y; if (p < p_lo || p > p_hi) continue; }
            const int got = fp_mul_apply<S, true>(mode == 2 ? fmaxf(__int2float_rn(x), xs) : __int2float_rn(x), yf, qq) - fp_kbits<S>();
            const int want = mul_apply<int32_t, false>(x, y, q);
            bad += got != want; ++seen;
        }
    }
    __shared__ unsigned long long sb[256], ss[256];
    sb[threadIdx.x] = bad; ss[threadIdx.x] = seen;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) { sb[threadIdx.x] += sb[threadIdx.x + o]; ss[threadIdx.x] += ss[threadIdx.x + o]; } __syncthreads(); }
    if (threadIdx.x == 0) { atomicAdd(&counts[0], ss[0]); atomicAdd(&counts[1], sb[0]); }
}

__global__ void __launch_bounds__(256) k_selftest_fadd(long long xlo, unsigned long long nx, long long ylo, unsigned long long ny,
                                                       const __grid_constant__ AddStage st, unsigned long long *counts)
{
    unsigned long long bad = 0, seen = 0;
    for (unsigned long long iy = blockIdx.x; iy < ny; iy += gridDim.x) {
        const int y = (int)(ylo + (long long)iy);
        const float yf = __int2float_rn(y);
        for (unsigned long long ix = threadIdx.x; ix < nx; ix += blockDim.x) {
            const int x = (int)(xlo + (long long)ix);
            const float r = fadd_stage(__int2float_rn(x), yf, st.q, st.sub != 0);
            const int want = add_apply<int32_t, false>(x, y, st);
            bad += (r != (float)want) || (__float2int_rn(r) != want); ++seen;
        }
    }
    __shared__ unsigned long long sb[256], ss[256];
    sb[threadIdx.x] = bad; ss[threadIdx.x] = seen;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) { sb[threadIdx.x] += sb[threadIdx.x + o]; ss[threadIdx.x] += ss[threadIdx.x + o]; } __syncthreads(); }
    if (threadIdx.x == 0) { atomicAdd(&counts[0], ss[0]); atomicAdd(&counts[1], sb[0]); }
}

static int launch_selftest_fpmul(int S, const Ival &x, const Ival &y, const QStage &q, int mode, long long p_lo, long long p_hi,
                                 unsigned long long *counts, cudaStream_t st, float xs = 0.f, float ys = 0.f)
{
    const unsigned long long nx = (unsigned long long)(x.hi - x.lo + 1), ny = (unsigned long long)(y.hi - y.lo + 1);
    const unsigned grid = (unsigned)std::min<unsigned long long>(ny, 148ull * 32);
#define QB_ST(SV) case SV: k_selftest_fpmul<SV><<<grid, 256, 0, st>>>((long long)x.lo, nx, (long long)y.lo, ny, q, mode, p_lo, p_hi, xs, ys, counts); break;
    switch (S) {
        QB_ST(1) QB_ST(2) QB_ST(3) QB_ST(4) QB_ST(5) QB_ST(6) QB_ST(7) QB_ST(8) QB_ST(9) QB_ST(10) QB_ST(11) QB_ST(12) QB_ST(13) QB_ST(14) QB_ST(15)
        QB_ST(16) QB_ST(17) QB_ST(18) QB_ST(19) QB_ST(20) QB_ST(21) QB_ST(22) QB_ST(23) QB_ST(24) QB_ST(25) QB_ST(26) QB_ST(27) QB_ST(28) QB_ST(29) QB_ST(30)
    default: return fail(QB_EINVAL, "FP product shift out of range");
    }
#undef QB_ST
    return post_launch("k_selftest_fpmul");
}

extern "C" int qb200_selftest_fp(int what, const void *plan, uint64_t *checked, uint64_t *mismatches, const qb_exec *exec)
{
    if (!plan || !checked || !mismatches) return fail(QB_EINVAL, "null argument");
    *checked = 0; *mismatches = 0;
    Ctx cx; int rc = cx.init(exec); if (rc) return rc;
    if (cx.host) { /* buffers are internal: the memory kind of the caller does not matter */ }
    unsigned long long *counts;
    if ((rc = cx.alloc((void **)&counts, 16))) return rc;
    CUDA_TRY(cudaMemsetAsync(counts, 0, 16, cx.stream));
    auto sweepable = [](const Ival &r) { return r.hi - r.lo < ((i128)1 << 24) && r.lo >= -((i128)1 << 24) && r.hi <= ((i128)1 << 24); };
    if (what == 0) {
        const qb_gemul_plan &p = *(const qb_gemul_plan *)plan;
        auto ce = get_gemul(p);
        if (ce->an.status) return fail(ce->an.status, ce->an.err);
        if (ce->bits == 32 && gemul_fp_mul(p, ce->prog)) {
            const Ival ra = type_range(p.a), rb = type_range(p.b);
            if (!sweepable(ra) || !sweepable(rb)) return fail(QB_EINVAL, "operand range too large to sweep");
            long long p_hi, p_lo;
            gemul_clamp_thresholds(ce->prog, &p_hi, &p_lo);
            if ((rc = launch_selftest_fpmul(ce->prog.mul.shift, ra, rb, ce->prog.mul, 0, 0, 0, counts, cx.stream))) return rc;
            const QStage &m = ce->prog.mul;
            if ((m.need_lo || m.need_hi) && (m.of == SAT_TCPL || m.of == SAT_SMGN))
                if ((rc = launch_selftest_fpmul(m.shift, ra, rb, m, 1, p_lo, p_hi, counts, cx.stream))) return rc;
            const FpSubst sub = fp_subst(true, m.shift, m, (long long)ra.lo, (long long)ra.hi, (long long)rb.lo, (long long)rb.hi);
            if (sub.ok && (rc = launch_selftest_fpmul(m.shift, ra, rb, m, 2, 0, 0, counts, cx.stream, sub.a_sub, sub.b_sub))) return rc;
        }
    } else if (what == 1) {
        const qb_cgemul_plan &p = *(const qb_cgemul_plan *)plan;
        auto ce = get_cgemul(p);
        if (ce->an.status) return fail(ce->an.status, ce->an.err);
        if (ce->bits == 32 && p.product == QB_CPROD_TF) {
            const CgemulFp f = cgemul_fp(p, ce->prog);
            const Ival a = type_range(p.a_re), b = type_range(p.a_im), c = type_range(p.b_re), d = type_range(p.b_im);
            const TfProg &tf = ce->prog.tf;
            for (const Ival *r : { &a, &b, &c, &d, &f.ab, &f.cd, &f.ba }) if (!sweepable(*r)) return fail(QB_EINVAL, "operand range too large to sweep");
            auto fadd = [&](const Ival &x, const Ival &y, const AddStage &st) -> int {
                const unsigned long long nx = (unsigned long long)(x.hi - x.lo + 1), ny = (unsigned long long)(y.hi - y.lo + 1);
                k_selftest_fadd<<<(unsigned)std::min<unsigned long long>(ny, 148ull * 32), 256, 0, cx.stream>>>((long long)x.lo, nx, (long long)y.lo, ny, st, counts);
                return post_launch("k_selftest_fadd");
            };
            if (f.fadd_ab && (rc = fadd(a, b, tf.ab))) return rc;                    // a + b  (QuBLAS.h:3524)
            if (f.fadd_ba && (rc = fadd(b, a, tf.ba))) return rc;                    // b - a  (QuBLAS.h:3526)
            if (f.fa && (rc = launch_selftest_fpmul(tf.A.shift, f.ab, c, tf.A, 0, 0, 0, counts, cx.stream))) return rc;   // (a+b) * c
            if (f.fb && (rc = launch_selftest_fpmul(tf.B.shift, f.cd, b, tf.B, 0, 0, 0, counts, cx.stream))) return rc;   // (c+d) * b
            if (f.fc && (rc = launch_selftest_fpmul(tf.C.shift, f.ba, d, tf.C, 0, 0, 0, counts, cx.stream))) return rc;   // (b-a) * d
        }
    } else {
        return fail(QB_EINVAL, "unknown plan kind");
    }
    unsigned long long h[2];
    CUDA_TRY(cudaMemcpyAsync(h, counts, 16, cudaMemcpyDeviceToHost, cx.stream));
    CUDA_TRY(cudaStreamSynchronize(cx.stream));
    *checked = h[0]; *mismatches = h[1];
    return cx.finish();
}

// ---------------------------------------------------------------------------------------------
// memory helpers for the host bindings (pinned host buffers, device-resident batches)
// ---------------------------------------------------------------------------------------------
extern "C" int qb200_host_alloc(void **ptr, size_t bytes, uint32_t flags)
{
    if (!ptr) return fail(QB_EINVAL, "null argument");
    *ptr = nullptr;
    if (flags & ~(uint32_t)QB_ALLOC_WRITE_COMBINED) return fail(QB_EINVAL, "unknown allocation flags");
    Ctx cx; int rc = cx.init(nullptr); if (rc) return rc;
    unsigned f = cudaHostAllocPortable | ((flags & QB_ALLOC_WRITE_COMBINED) ? cudaHostAllocWriteCombined : 0);
    cudaError_t e = cudaHostAlloc(ptr, bytes ? bytes : 16, f);
    if (e != cudaSuccess) { cudaGetLastError(); *ptr = nullptr; return fail(e == cudaErrorMemoryAllocation ? QB_ENOMEM : QB_ECUDA, std::string("cudaHostAlloc: ") + cudaGetErrorString(e)); }
    return QB_OK;
}
extern "C" int qb200_host_free(void *ptr)
{
    if (!ptr) return QB_OK;
    CUDA_TRY(cudaFreeHost(ptr));
    return QB_OK;
}
extern "C" int qb200_device_alloc(void **ptr, size_t bytes, int device)
{
    if (!ptr) return fail(QB_EINVAL, "null argument");
    *ptr = nullptr;
    qb_exec ex = { QB_MEM_DEVICE, device, nullptr, 0, 0 };
    Ctx cx; int rc = cx.init(&ex); if (rc) return rc;
    cudaError_t e = cudaMalloc(ptr, bytes ? bytes : 16);
    if (e != cudaSuccess) { cudaGetLastError(); *ptr = nullptr; return fail(e == cudaErrorMemoryAllocation ? QB_ENOMEM : QB_ECUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e)); }
    return QB_OK;
}
extern "C" int qb200_device_free(void *ptr, int device)
{
    if (!ptr) return QB_OK;
    if (device >= 0) CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaFree(ptr));
    return QB_OK;
}
// dst / src may be host or device memory (unified addressing decides); synchronous on return
extern "C" int qb200_memcpy(void *dst, const void *src, size_t bytes, int device)
{
    if (!dst || !src) return fail(QB_EINVAL, "null argument");
    if (!bytes) return QB_OK;
    qb_exec ex = { QB_MEM_DEVICE, device, nullptr, 0, 0 };
    Ctx cx; int rc = cx.init(&ex); if (rc) return rc;
    CUDA_TRY(cudaMemcpy(dst, src, bytes, cudaMemcpyDefault));
    return QB_OK;
}

// ---------------------------------------------------------------------------------------------
// float64 reference GEMM of the DEQUANTIZED operands (the yardstick of an accuracy sweep):
// C64[i,j] = sum_k toDouble(opA(i,k)) * toDouble(opB(k,j)), k ascending, one FMA per term (deterministic).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_gemul_f64(const void *__restrict__ A, const void *__restrict__ B, double *__restrict__ C, uint64_t batch,
                                                   const __grid_constant__ GemulShape S, int frac_a, int frac_b)
{
    const uint64_t per = (uint64_t)S.M * S.N;
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= batch * per) return;
    const uint64_t inst = t / per, ij = t - inst * per;
    const uint32_t j = (uint32_t)(ij / S.M), i = (uint32_t)(ij - (uint64_t)j * S.M);
    const uint64_t oa = inst * (uint64_t)S.stride_a, ob = inst * (uint64_t)S.stride_b;
    double acc = 0.0;
    for (uint64_t k = 0; k < S.K; ++k) {
        const uint64_t ia = S.trans_a ? (k + (uint64_t)i * S.K) : (i + k * S.M);
        const uint64_t ib = S.trans_b ? (j + k * S.N) : (k + (uint64_t)j * S.K);
        acc = fma((double)load_word<int64_t>(A, oa + ia, S.a_wb), (double)load_word<int64_t>(B, ob + ib, S.b_wb), acc);
    }
    C[inst * (uint64_t)S.stride_c + ij] = acc * pow2_neg(frac_a + frac_b);         // raw products are exact integers below 2^53 per term
}

extern "C" int qb200_gemul_f64_batched(const qb_gemul_plan *plan, const void *A, const void *B, double *C, uint64_t batch, const qb_exec *exec)
{
    if (!plan || !A || !B || !C) return fail(QB_EINVAL, "null argument");
    auto ce = get_gemul(*plan);
    if (ce->an.status) return fail(ce->an.status, ce->an.err);
    const GemulProg &gp = ce->prog;
    if (width(plan->a) + width(plan->b) > 53) return fail(QB_EDOMAIN, "operand products do not fit a float64 mantissa");
    if (plan->stride_a < 0 || plan->stride_b < 0 || plan->stride_c < (int64_t)((uint64_t)plan->M * plan->N)) return fail(QB_EINVAL, "bad batch strides");
    if (batch == 0) return QB_OK;
    Ctx cx; int rc = cx.init(exec); if (rc) return rc;
    const uint64_t ea = (uint64_t)plan->M * plan->K, eb = (uint64_t)plan->K * plan->N, ec = (uint64_t)plan->M * plan->N;
    const size_t bytes_a = span_elems(batch, plan->stride_a, ea) * gp.a_wb, bytes_b = span_elems(batch, plan->stride_b, eb) * gp.b_wb,
                 bytes_c = span_elems(batch, plan->stride_c, ec) * sizeof(double);
    const void *dA, *dB; void *dC;
    if ((rc = cx.in(A, bytes_a, &dA)) || (rc = cx.in(B, bytes_b, &dB)) || (rc = cx.out(C, bytes_c, &dC))) return rc;
    GemulShape S = shape_of(plan->M, plan->N, plan->K, plan->trans_a, plan->trans_b, gp.a_wb, gp.b_wb, 8, plan->stride_a, plan->stride_b, plan->stride_c);
    const uint64_t blocks = (batch * ec + 127) / 128;
    if (blocks > 0x7fffffffull) return fail(QB_EINVAL, "too many outputs for one launch");
    k_gemul_f64<<<(unsigned)blocks, 128, 0, cx.stream>>>(dA, dB, (double *)dC, batch, S, plan->a.frac_bits, plan->b.frac_bits);
    if ((rc = post_launch("k_gemul_f64"))) return rc;
    if ((rc = cx.out_done(C, bytes_c, dC))) return rc;
    return cx.finish();
}

// ---------------------------------------------------------------------------------------------
// Accuracy of a quantized batch in ONE pass (the Monte-Carlo sweep of north_star: Qgemul -> error against float64): the
// float64 reference of every instance is formed in registers and compared with the quantized result on the spot, so neither
// the batch x M x N doubles nor a second pass over them exist.  One CTA per instance at a time: the operands are converted to
// double ONCE into shared memory (A as [k][i], B as [k][j]), every thread owns 4 x 4 outputs and runs the same k-ascending FMA
// chain per output as k_gemul_f64 (bit-identical reference values); the sums are reduced in a fixed order (fixed grid, fixed
// instance -> CTA mapping, fixed tree), i.e. reproducible run to run.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gemul_err_f64(const void *__restrict__ A, const void *__restrict__ B, const void *__restrict__ C, uint64_t batch,
                                                       const __grid_constant__ GemulShape S, int frac_a, int frac_b, int frac_c, double *__restrict__ part)
{
    extern __shared__ __align__(16) double err_sm[];          // 16-byte loads of four consecutive rows / columns
    double *As = err_sm, *Bt = err_sm + (size_t)S.K * S.M;
    __shared__ double s0[256], s1[256], s2[256];
    const uint32_t M = S.M, N = S.N, K = S.K, tid = threadIdx.x;
    const uint32_t tm = M / 4, ntiles = tm * (N / 4);
    const double sc = pow2_neg(frac_a + frac_b), scc = pow2_neg(frac_c);
    double se = 0, sr = 0, mx = 0;
    for (uint64_t inst = blockIdx.x; inst < batch; inst += gridDim.x) {
        __syncthreads();                                        // the previous instance's tiles are done with shared memory
        const uint64_t oa = inst * (uint64_t)S.stride_a, ob = inst * (uint64_t)S.stride_b;
        for (uint32_t r = tid; r < M * K; r += blockDim.x) {    // stored order (coalesced); As[k][i]
            const double v = (double)load_word<int64_t>(A, oa + r, S.a_wb);
            if (S.trans_a) { const uint32_t i = r / K, k = r - i * K; As[k * M + i] = v; } else As[r] = v;
        }
        for (uint32_t r = tid; r < K * N; r += blockDim.x) {    // Bt[k][j]
            const double v = (double)load_word<int64_t>(B, ob + r, S.b_wb);
            if (S.trans_b) Bt[r] = v; else { const uint32_t j = r / K, k = r - j * K; Bt[k * N + j] = v; }
        }
        __syncthreads();
        for (uint32_t t = tid; t < ntiles; t += blockDim.x) {
            const uint32_t i0 = (t % tm) * 4, j0 = (t / tm) * 4;
            double acc[4][4];
#pragma unroll
            for (int ii = 0; ii < 4; ii++)
#pragma unroll
                for (int jj = 0; jj < 4; jj++) acc[ii][jj] = 0.0;
#pragma unroll 4
            for (uint32_t k = 0; k < K; ++k) {
                const double2 a01 = *(const double2 *)(As + k * M + i0), a23 = *(const double2 *)(As + k * M + i0 + 2);
                const double2 b01 = *(const double2 *)(Bt + k * N + j0), b23 = *(const double2 *)(Bt + k * N + j0 + 2);
                const double a[4] = { a01.x, a01.y, a23.x, a23.y }, b[4] = { b01.x, b01.y, b23.x, b23.y };
#pragma unroll
                for (int ii = 0; ii < 4; ii++)
#pragma unroll
                    for (int jj = 0; jj < 4; jj++) acc[ii][jj] = fma(a[ii], b[jj], acc[ii][jj]);
            }
            const uint64_t oc = inst * (uint64_t)S.stride_c;
#pragma unroll
            for (int jj = 0; jj < 4; jj++)
#pragma unroll
                for (int ii = 0; ii < 4; ii++) {
                    const double ref = acc[ii][jj] * sc;
                    const double d = (double)load_word<int64_t>(C, oc + (uint64_t)(j0 + jj) * M + i0 + ii, S.c_wb) * scc - ref;
                    se += d * d; sr += ref * ref; mx = fmax(mx, fabs(d));
                }
        }
    }
    s0[tid] = se; s1[tid] = sr; s2[tid] = mx;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)tid < o) { s0[tid] += s0[tid + o]; s1[tid] += s1[tid + o]; s2[tid] = fmax(s2[tid], s2[tid + o]); }
        __syncthreads();
    }
    if (tid == 0) { part[3 * blockIdx.x] = s0[0]; part[3 * blockIdx.x + 1] = s1[0]; part[3 * blockIdx.x + 2] = s2[0]; }
}

extern "C" int qb200_gemul_error_batched(const qb_gemul_plan *plan, const void *A, const void *B, const void *C, uint64_t batch,
                                         qb_err_stats *stats, const qb_exec *exec)
{
    if (!plan || !A || !B || !C || !stats) return fail(QB_EINVAL, "null argument");
    if (wants_all_devices(exec)) return fail(QB_EINVAL, "error statistics run on one device (reduce per device and add the sums)");
    auto ce = get_gemul(*plan);
    if (ce->an.status) return fail(ce->an.status, ce->an.err);
    const GemulProg &gp = ce->prog;
    if (width(plan->a) + width(plan->b) > 53) return fail(QB_EDOMAIN, "operand products do not fit a float64 mantissa");
    if (plan->stride_a < 0 || plan->stride_b < 0 || plan->stride_c < (int64_t)((uint64_t)plan->M * plan->N)) return fail(QB_EINVAL, "bad batch strides");
    const uint64_t ea = (uint64_t)plan->M * plan->K, eb = (uint64_t)plan->K * plan->N, ec = (uint64_t)plan->M * plan->N;
    memset(stats, 0, sizeof(*stats));
    stats->n = batch * ec;
    if (batch == 0) return QB_OK;
    Ctx cx; int rc = cx.init(exec); if (rc) return rc;
    const size_t bytes_a = span_elems(batch, plan->stride_a, ea) * gp.a_wb, bytes_b = span_elems(batch, plan->stride_b, eb) * gp.b_wb,
                 bytes_c = span_elems(batch, plan->stride_c, ec) * gp.c_wb;
    const void *dA, *dB, *dC;
    if ((rc = cx.in(A, bytes_a, &dA)) || (rc = cx.in(B, bytes_b, &dB)) || (rc = cx.in(C, bytes_c, &dC))) return rc;
    GemulShape S = shape_of(plan->M, plan->N, plan->K, plan->trans_a, plan->trans_b, gp.a_wb, gp.b_wb, gp.c_wb, plan->stride_a, plan->stride_b, plan->stride_c);
    const size_t smem = (ea + eb) * sizeof(double);
    void *part, *res;
    if ((rc = cx.alloc(&res, 3 * 8))) return rc;
    int nblocks = 0;
    if (plan->M % 4 == 0 && plan->N % 4 == 0 && smem <= 200 * 1024) {
        int dev = 0, sms = 0, per_sm = 0;
        CUDA_TRY(cudaGetDevice(&dev));
        CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        CUDA_TRY(cudaFuncSetAttribute(k_gemul_err_f64, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gemul_err_f64, 256, smem));
        const uint64_t cap = (uint64_t)sms * (per_sm > 0 ? per_sm : 1);
        nblocks = (int)(batch < cap ? batch : cap);
        if ((rc = cx.alloc(&part, 3 * (size_t)nblocks * 8))) return rc;
        k_gemul_err_f64<<<nblocks, 256, smem, cx.stream>>>(dA, dB, dC, batch, S, plan->a.frac_bits, plan->b.frac_bits, plan->c.frac_bits, (double *)part);
        if ((rc = post_launch("k_gemul_err_f64"))) return rc;
    } else {
        // other shapes: the float64 reference through the one-thread-per-output kernel, then the streaming comparison
        void *ref;
        nblocks = 1184;
        if (plan->stride_c != (int64_t)ec) return fail(QB_EINVAL, "padded C batches need M, N multiples of 4 here");
        if ((rc = cx.alloc(&ref, batch * ec * sizeof(double))) || (rc = cx.alloc(&part, 3 * (size_t)nblocks * 8))) return rc;
        GemulShape S64 = S; S64.c_wb = 8;
        const uint64_t blocks = (batch * ec + 127) / 128;
        if (blocks > 0x7fffffffull) return fail(QB_EINVAL, "too many outputs for one launch");
        k_gemul_f64<<<(unsigned)blocks, 128, 0, cx.stream>>>(dA, dB, (double *)ref, batch, S64, plan->a.frac_bits, plan->b.frac_bits);
        if ((rc = post_launch("k_gemul_f64"))) return rc;
        k_err_stats<<<nblocks, 256, 0, cx.stream>>>(dC, (const double *)ref, batch * ec, gp.c_wb, plan->c.frac_bits, (double *)part);
        if ((rc = post_launch("k_err_stats"))) return rc;
    }
    k_err_final<<<1, 256, 0, cx.stream>>>((const double *)part, nblocks, (double *)res);
    if ((rc = post_launch("k_err_final"))) return rc;
    double h[3];
    CUDA_TRY(cudaMemcpyAsync(h, res, sizeof(h), cudaMemcpyDeviceToHost, cx.stream));
    CUDA_TRY(cudaStreamSynchronize(cx.stream));
    stats->sum_sq_err = h[0]; stats->sum_sq_ref = h[1]; stats->max_abs_err = h[2];
    return cx.finish();
}

// ---------------------------------------------------------------------------------------------
// device-resident shards on several GPUs in one call, and the optional output all-gather over NVLink
// ---------------------------------------------------------------------------------------------
extern "C" int qb200_gemul_sharded(const qb_gemul_plan *plan, const qb_shard *shards, uint32_t n_shards, const qb_exec *exec)
{
    if (!plan || (!shards && n_shards)) return fail(QB_EINVAL, "null argument");
    if (exec && exec->stream) return fail(QB_EINVAL, "sharded calls run on the library's own streams (exec.stream must be NULL)");
    if (n_shards == 0) return QB_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    std::vector<int> rcs(n_shards, QB_OK);
    std::vector<std::string> errs(n_shards);
    std::vector<std::thread> th;
    for (uint32_t r = 0; r < n_shards; r++) {
        th.emplace_back([&, r] {
            const qb_shard &sh = shards[r];
            qb_exec ex = { exec ? exec->mem_kind : (int32_t)QB_MEM_DEVICE, sh.device, nullptr, exec ? exec->flags : 0u, 0 };
            int rc = sh.batch ? qb200_gemul_batched(plan, sh.A, sh.B, sh.C, sh.batch, &ex) : QB_OK;
            if (!rc && ex.mem_kind == QB_MEM_DEVICE) {                       // device-memory calls are asynchronous: complete the shard before returning
                cudaError_t e = cudaStreamSynchronize(nullptr);
                if (e != cudaSuccess) rc = fail(QB_ECUDA, std::string("cudaStreamSynchronize: ") + cudaGetErrorString(e));
            }
            rcs[r] = rc;
            if (rc) errs[r] = g_err;
        });
    }
    for (auto &t : th) t.join();
    cudaSetDevice(prev);
    for (uint32_t r = 0; r < n_shards; r++) if (rcs[r]) return fail(rcs[r], "shard " + std::to_string(r) + " (device " + std::to_string(shards[r].device) + "): " + errs[r]);
    return QB_OK;
}

// Every device ends up with every shard: direct peer copies over NVLink / NVSwitch (no host staging, no compute-path
// dependency: it is optional and separate from the kernels).  Shard i is `rows` rows of width[i] bytes with pitch
// src_pitch[i] on devices[i]; it lands at byte offset dst_off[i] of every row of dst[j] (pitch dst_pitch) on every device j.
// rows == 1 is the contiguous case (batch shards); rows = N with column-major M x N matrices gathers M-row panels.
static int enable_peers(const int32_t *devices, uint32_t n)
{
    int prev = 0;
    cudaGetDevice(&prev);
    for (uint32_t i = 0; i < n; i++) {
        CUDA_TRY(cudaSetDevice(devices[i]));
        for (uint32_t j = 0; j < n; j++) {
            if (devices[j] == devices[i]) continue;
            int can = 0;
            CUDA_TRY(cudaDeviceCanAccessPeer(&can, devices[i], devices[j]));
            if (!can) { cudaSetDevice(prev); return fail(QB_ECUDA, "no peer access between devices " + std::to_string(devices[i]) + " and " + std::to_string(devices[j])); }
            cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaSetDevice(prev); return fail(QB_ECUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)); }
            cudaGetLastError();
        }
    }
    cudaSetDevice(prev);
    return QB_OK;
}

// GEMM -> all-gather in ONE kernel per GPU (the optional output all-gather of SURVEY.md 8(e), fused): shard r computes the
// M-row panel r of a (n * M) x N column-major C and its tensor-core epilogue stores every element into the C of EVERY
// device -- its own through local stores, the others' through peer stores over NVLink / NVSwitch (128 contiguous bytes per
// warp and column), overlapped tile by tile with the MMAs of the next tile through the second TMEM accumulator.  No
// separate copy pass, no staging of the panel.  plan describes ONE panel (plan->M rows); c_full[j] is the whole C on
// shards[j].device (leading dimension n * M); shards[r].C is ignored.  Tensor-core plans only (lossless int8-range
// accumulation, M a multiple of 256); n <= 8.  Complete on return.
extern "C" int qb200_gemul_panels_gathered(const qb_gemul_plan *plan, const qb_shard *shards, uint32_t n, void *const *c_full, const qb_exec *exec)
{
    if (!plan || !shards || !c_full) return fail(QB_EINVAL, "null argument");
    if (n == 0) return QB_OK;
    if (n > 8) return fail(QB_EINVAL, "at most 8 panels");
    if (exec && (exec->stream || exec->mem_kind != QB_MEM_DEVICE)) return fail(QB_EINVAL, "device-resident panels on the library's own streams");
    int32_t devs[8];
    for (uint32_t r = 0; r < n; r++) { devs[r] = shards[r].device; if (!shards[r].A || !shards[r].B || !c_full[r]) return fail(QB_EINVAL, "null shard pointer"); }
    int rc = enable_peers(devs, n); if (rc) return rc;
    int prev = 0;
    cudaGetDevice(&prev);
    std::vector<int> rcs(n, QB_OK);
    std::vector<std::string> errs(n);
    std::vector<std::thread> th;
    for (uint32_t r = 0; r < n; r++) {
        th.emplace_back([&, r] {
            I8FanOut fan;
            fan.n = n; fan.row0 = r * plan->M; fan.ldc = (uint64_t)n * plan->M;
            for (uint32_t j = 0; j < 8; j++) fan.out[j] = j < n ? c_full[j] : nullptr;
            qb_exec ex = { (int32_t)QB_MEM_DEVICE, shards[r].device, nullptr, exec ? exec->flags : 0u, 0 };
            tl_fanout = &fan;
            int e = qb200_gemul_batched(plan, shards[r].A, shards[r].B, c_full[r], 1, &ex);
            tl_fanout = nullptr;
            if (!e) {
                cudaError_t ce = cudaStreamSynchronize(nullptr);
                if (ce != cudaSuccess) e = fail(QB_ECUDA, std::string("cudaStreamSynchronize: ") + cudaGetErrorString(ce));
            }
            rcs[r] = e;
            if (e) errs[r] = g_err;
        });
    }
    for (auto &t : th) t.join();
    cudaSetDevice(prev);
    for (uint32_t r = 0; r < n; r++) if (rcs[r]) return fail(rcs[r], "panel " + std::to_string(r) + " (device " + std::to_string(shards[r].device) + "): " + errs[r]);
    return QB_OK;
}

extern "C" int qb200_allgather(void *const *dst, const void *const *src, const int32_t *devices, uint32_t n, uint64_t rows, uint64_t dst_pitch,
                               const uint64_t *dst_off, const uint64_t *src_pitch, const uint64_t *width, double *seconds)
{
    if (!dst || !src || !devices || !dst_off || !width || (rows > 1 && !src_pitch)) return fail(QB_EINVAL, "null argument");
    if (n == 0 || rows == 0) return QB_OK;
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    for (uint32_t i = 0; i < n; i++) if (devices[i] < 0 || devices[i] >= ndev) return fail(QB_EINVAL, "bad device ordinal");
    int prev = 0;
    cudaGetDevice(&prev);
    // peer access in both directions between every pair (a copy without it would be staged through the host)
    for (uint32_t i = 0; i < n; i++) {
        CUDA_TRY(cudaSetDevice(devices[i]));
        for (uint32_t j = 0; j < n; j++) {
            if (devices[j] == devices[i]) continue;
            int can = 0;
            CUDA_TRY(cudaDeviceCanAccessPeer(&can, devices[i], devices[j]));
            if (!can) { cudaSetDevice(prev); return fail(QB_ECUDA, "no peer access between devices " + std::to_string(devices[i]) + " and " + std::to_string(devices[j])); }
            cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaSetDevice(prev); return fail(QB_ECUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)); }
            cudaGetLastError();
        }
    }
    std::vector<cudaStream_t> st(n);
    std::vector<cudaEvent_t> e0(n), e1(n);
    for (uint32_t i = 0; i < n; i++) {
        CUDA_TRY(cudaSetDevice(devices[i]));
        CUDA_TRY(cudaDeviceSynchronize());                                  // the shards are complete before anybody reads them
        CUDA_TRY(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreate(&e0[i])); CUDA_TRY(cudaEventCreate(&e1[i]));
    }
    int rc = QB_OK;
    for (uint32_t i = 0; i < n && !rc; i++) {                               // source device i pushes its shard: step s goes to (i + s) mod n
        cudaSetDevice(devices[i]);
        cudaEventRecord(e0[i], st[i]);
        for (uint32_t s = 0; s < n && !rc; s++) {
            const uint32_t j = (i + s) % n;
            char *d = (char *)dst[j] + dst_off[i];
            const uint64_t sp = rows > 1 ? src_pitch[i] : width[i];
            if ((const void *)d == src[i] && (rows == 1 || sp == dst_pitch)) continue;        // already in place on its own device
            cudaError_t e;
            if (rows == 1) e = cudaMemcpyPeerAsync(d, devices[j], src[i], devices[i], width[i], st[i]);
            else e = cudaMemcpy2DAsync(d, dst_pitch, src[i], sp, width[i], rows, cudaMemcpyDefault, st[i]);
            if (e != cudaSuccess) rc = fail(QB_ECUDA, std::string("peer copy: ") + cudaGetErrorString(e));
        }
        cudaEventRecord(e1[i], st[i]);
    }
    double worst = 0;
    for (uint32_t i = 0; i < n; i++) {
        cudaSetDevice(devices[i]);
        cudaError_t e = cudaStreamSynchronize(st[i]);
        if (e != cudaSuccess && !rc) rc = fail(QB_ECUDA, std::string("all-gather: ") + cudaGetErrorString(e));
        float ms = 0;
        if (!rc && cudaEventElapsedTime(&ms, e0[i], e1[i]) == cudaSuccess && ms * 1e-3 > worst) worst = ms * 1e-3;
        cudaStreamDestroy(st[i]); cudaEventDestroy(e0[i]); cudaEventDestroy(e1[i]);
    }
    cudaSetDevice(prev);
    if (seconds) *seconds = worst;
    return rc;
}
