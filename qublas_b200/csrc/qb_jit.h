// qb_jit.h — run-time specialisation of the kernel templates with NVRTC.
//
// The AOT registry (qb_fast.cu) covers the BASELINE configurations; every other plan whose shape
// fits a specialised kernel template gets its own instantiation at first use: qb_codegen prints the
// compiled program as a tag struct, NVRTC compiles `kernel<Tag>` for sm_100a, the cubin is loaded
// through the driver API and cached with the compiled plan.  This is the literal form of "the host
// side lowers each Qu type's compile-time parameters into kernel template arguments".
// If NVRTC is not available the library falls back to the runtime-descriptor CUDA kernels
// (still the GPU — never a CPU path).  QB200_NO_JIT=1 disables it.
#pragma once
#include <cuda_runtime.h>
#include <memory>
#include <string>
#include "qb_plan.h"

namespace qb {

struct JitKernel {
    void *module = nullptr;        // CUmodule
    void *func[4] = { nullptr, nullptr, nullptr, nullptr };   // CUfunction(s)
    int npass = 0;                                             // > 0: any-length reduction, one entry point per pass
    bool variadic = false;                                     // the passes reduce whole chunks only; k_reduce_var_finish completes the tree
    struct { uint64_t n_in, n_out; } pass[4] = {};
    unsigned threads = 0, smem = 0;
    int blocks_per_sm = 1, sms = 0;
    unsigned group = 1;                     // tile template: instances one CTA works on at a time
    unsigned outputs_per_unit = 0;          // > 0: thread-per-output kernel, grid = ceil(batch * outputs_per_unit / threads)
    std::string name, log;
    bool ok = false;
};

bool jit_enabled();
// Each returns nullptr when the plan does not fit the template or when compilation fails.
// epi: ANUS polynomial fused into the store of every result element (nullptr: none)
// pro_a / pro_b: ANUS polynomial fused into the loads of operand A / B (nullptr: none)
std::shared_ptr<JitKernel> jit_gemul(const qb_gemul_plan &p, const GemulProg &g, int arith_bits, const PolyProg *epi = nullptr, const PolyProg *pro_a = nullptr,
                                     const PolyProg *pro_b = nullptr);
std::shared_ptr<JitKernel> jit_gemul_thread(const qb_gemul_plan &p, const GemulProg &g, int arith_bits);   // thread-per-output template only
std::shared_ptr<JitKernel> jit_poly(const PolyProg &prog, int xw, int ow, int arith_bits);
std::shared_ptr<JitKernel> jit_cgemul(const qb_cgemul_plan &p, const CGemulProg &g, int arith_bits);
std::shared_ptr<JitKernel> jit_reduce(const qb_reduce_plan &p, const ReduceProg &r, int arith_bits);

// element-wise templates (qb_fast_ew.cuh): kind >= 0 complex product (tf / bcm), kind < 0 real op (0 mul, 1 add, 2 sub, 3 convert)
std::shared_ptr<JitKernel> jit_ew(int kind, const TfProg *tf, const BcmProg *bcm, int op, const AddStage *stage, int xw, int yw, int ow, int arith_bits);
int jit_launch_ew(const JitKernel &k, const void *x, const void *y, void *out, uint64_t n, cudaStream_t s);

// compile-only mode for the calling thread: jit_* build the source and run NVRTC, nothing is loaded (no GPU needed)
void jit_set_compile_only(bool on);
int jit_compile_only_result(std::string &log);   // 1 = no template applied, 0 = cubin produced, -1 = NVRTC failed

int jit_check_gemul(const qb_gemul_plan &p, const GemulProg &g, int arith_bits, std::string &log);

// launches; return 0 or a CUresult / cudaError code
int jit_launch_gemul(const JitKernel &k, const void *A, const void *B, void *C, uint64_t batch, cudaStream_t s);
int jit_launch_reduce(const JitKernel &k, unsigned logn, const void *x, void *out, void *scratch, uint64_t batch, cudaStream_t s);
// any-length variant: scratch must hold 2 * batch * pass[0].n_out words (two ping-pong halves)
int jit_launch_reduce_any(const JitKernel &k, const void *x, void *out, void *scratch, uint64_t batch, cudaStream_t s);
// variadic form, length not a power of two: whole 256-node chunks per pass; roots_of[i] = where pass i wrote its batch * n_out roots
int jit_launch_reduce_var(const JitKernel &k, const void *x, void *roots, uint64_t batch, cudaStream_t s, const int32_t **roots_of);

}  // namespace qb
