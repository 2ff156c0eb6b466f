// e2e_c3.cpp — BASELINE config 3 driven end to end through the C++ drop-in API (include/QuBLAS.h), exactly as user code
// would write it: Qbatch containers in pinned host memory (the header's native narrow words: int16 operands, int32
// results), QgemulBatched<QgemulAddArgs<TypeList<...>>, QgemulMulArgs<...>>(C, A, B) per chunk.  Built by
// __graft_entry__.build() into qublas_b200/libqb200_e2e.so and called by bench.py for the `e2e` figure, so that the
// number is measured through the reference-facing call and not through a Python binding of the C ABI.
#include "QuBLAS.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>

using namespace QuBLAS;

namespace {
using in_t  = Qu<intBits<7>, fracBits<8>>;
using mul_t = Qu<intBits<14>, fracBits<8>, QuMode<RND::CONV>, OfMode<SAT::TCPL>>;
template <int l> using lv = Qu<intBits<15 + l>, fracBits<8 - l>, QuMode<RND::POS_INF>, OfMode<SAT::TCPL>>;
using levels = TypeList<lv<0>, lv<1>, lv<2>, lv<3>, lv<4>, lv<5>>;
using c_t = Qu<intBits<17>, fracBits<4>, QuMode<RND::POS_INF>, OfMode<SAT::TCPL>>;
using A_t = Qu<dim<64, 64>, in_t>;
using C_t = Qu<dim<64, 64>, c_t>;

template <Qwords W> struct Ring {
    static constexpr int SLOTS = 2;
    std::unique_ptr<Qbatch<A_t, W>> A[SLOTS], B[SLOTS];
    std::unique_ptr<Qbatch<C_t, W>> C[SLOTS];
    std::size_t chunk = 0;
    int device = -1;
    void init(std::size_t n, int dev, std::uint64_t first_instance)
    {
        chunk = n; device = dev;
        for (int s = 0; s < SLOTS; s++) {
            const char *wc = std::getenv("QB200_E2E_WC");                        // experiment: write-combined input batches
            const Qmem in_mem = (wc && wc[0] == '1') ? Qmem::write_combined : Qmem::pinned;
            A[s] = std::make_unique<Qbatch<A_t, W>>(n, in_mem); B[s] = std::make_unique<Qbatch<A_t, W>>(n, in_mem); C[s] = std::make_unique<Qbatch<C_t, W>>(n);
            if (A[s]->memory() == Qmem::pageable || C[s]->memory() == Qmem::pageable) throw std::runtime_error("pinned host memory unavailable");
            A[s]->synth(1, 0, first_instance + s * n); B[s]->synth(1, 1, first_instance + s * n);     // the bench's synthetic stream
        }
    }
    double step(std::size_t instances)
    {
        const auto t0 = std::chrono::steady_clock::now();
        std::size_t done = 0;
        for (int it = 0; done < instances; ++it) {
            const int s = it % SLOTS;
            // a short last chunk would need a second container; the bench uses multiples of the chunk
            QgemulBatched<QgemulAddArgs<levels>, QgemulMulArgs<mul_t>>(*C[s], *A[s], *B[s], device);
            done += chunk;
        }
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
};
Ring<Qwords::narrow> g_narrow;
Ring<Qwords::reference> g_ref;
char g_err[512];
}  // namespace

extern "C" {
const char *qbx_last_error(void) { return g_err; }

// words: 0 = the header's native narrow words (int16 operands), 1 = the reference's words (int32 everywhere)
int qbx_c3_init(int words, std::uint64_t chunk, int device, std::uint64_t first_instance)
{
    try { if (words) g_ref.init(chunk, device, first_instance); else g_narrow.init(chunk, device, first_instance); return 0; }
    catch (const std::exception &e) { std::snprintf(g_err, sizeof(g_err), "%s", e.what()); return -1; }
}
// one step = `instances` instances (rounded up to whole chunks) through QgemulBatched from host memory; returns seconds
int qbx_c3_step(int words, std::uint64_t instances, double *seconds, std::uint64_t *h2d_bytes, std::uint64_t *d2h_bytes)
{
    try {
        const std::size_t chunk = words ? g_ref.chunk : g_narrow.chunk;
        const std::uint64_t calls = (instances + chunk - 1) / chunk;
        *seconds = words ? g_ref.step(instances) : g_narrow.step(instances);
        *h2d_bytes = calls * (words ? g_ref.A[0]->bytes() + g_ref.B[0]->bytes() : g_narrow.A[0]->bytes() + g_narrow.B[0]->bytes());
        *d2h_bytes = calls * (words ? g_ref.C[0]->bytes() : g_narrow.C[0]->bytes());
        return 0;
    } catch (const std::exception &e) { std::snprintf(g_err, sizeof(g_err), "%s", e.what()); return -1; }
}
// the first `n_words` result words of slot 0 (instance first_instance ...), widened to int32, for the bench's equality check
int qbx_c3_result(int words, std::int32_t *out, std::uint64_t n_words)
{
    for (std::uint64_t i = 0; i < n_words; i++) out[i] = words ? (std::int32_t)g_ref.C[0]->raw()[i] : (std::int32_t)g_narrow.C[0]->raw()[i];
    return 0;
}
void qbx_c3_free(void)
{
    for (int s = 0; s < 2; s++) { g_narrow.A[s].reset(); g_narrow.B[s].reset(); g_narrow.C[s].reset(); g_ref.A[s].reset(); g_ref.B[s].reset(); g_ref.C[s].reset(); }
}

// The Monte-Carlo sweep of north_star, device-resident: inputs generated in HBM, Qgemul, float64 yardstick and error
// statistics on the GPU; 32 bytes come back.  Returns seconds for the whole sweep of `instances` instances in chunks.
int qbx_c3_sweep(std::uint64_t instances, std::uint64_t chunk, int device, std::uint64_t first_instance, double *seconds, double *sqnr_db, double *max_abs_err)
{
    try {
        Qbatch<A_t> A(chunk, Qmem::device, device), B(chunk, Qmem::device, device); Qbatch<C_t> C(chunk, Qmem::device, device);
        Qerror tot;
        const auto t0 = std::chrono::steady_clock::now();
        for (std::uint64_t done = 0; done < instances; done += chunk) {
            A.synth(1, 0, first_instance + done); B.synth(1, 1, first_instance + done);
            QgemulBatched<QgemulAddArgs<levels>, QgemulMulArgs<mul_t>>(C, A, B);
            Qerror e = QgemulError<QgemulAddArgs<levels>, QgemulMulArgs<mul_t>>(C, A, B);
            tot.sum_sq_err += e.sum_sq_err; tot.sum_sq_ref += e.sum_sq_ref; tot.n += e.n;
            if (e.max_abs_err > tot.max_abs_err) tot.max_abs_err = e.max_abs_err;
        }
        *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        *sqnr_db = tot.sqnr_db(); *max_abs_err = tot.max_abs_err;
        return 0;
    } catch (const std::exception &e) { std::snprintf(g_err, sizeof(g_err), "%s", e.what()); return -1; }
}
}
