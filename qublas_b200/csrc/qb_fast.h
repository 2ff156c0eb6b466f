// qb_fast.h — registry of compile-time-specialised kernels (qb_fast_*.cu).
//
// A specialised kernel is the same code as the runtime-descriptor kernel with the plan's program
// baked in as constants ("each Qu type's compile-time parameters lowered into kernel template
// arguments"): shifts and bounds are immediates, dead clamps vanish, the k-loop and the
// level-indexed stack are fully unrolled into registers.  The registry is keyed by the exact plan.
#pragma once
#include <cuda_runtime.h>
#include "../../include/qb200.h"
#include "qb_plan.h"

namespace qb {

void count_launch(int n);

struct FastGemul {
    const char *name;
    qb_gemul_plan plan;        // exact plan this kernel was specialised for (strides = dense)
    int (*launch)(const void *A, const void *B, void *C, uint64_t batch, cudaStream_t s);   // returns cudaError_t
};
struct FastCGemul {
    const char *name;
    qb_cgemul_plan plan;
    int (*launch)(const void *A, const void *B, void *C, uint64_t batch, cudaStream_t s);
};

struct FastReduce {
    const char *name;
    qb_reduce_plan plan;
    int (*launch)(const void *x, void *out, void *scratch, uint64_t batch, cudaStream_t s);      // scratch: batch * 4096 int32
};
const FastReduce *find_fast_reduce(const qb_reduce_plan &p, const ReduceProg &r, int arith_bits);
const FastGemul  *find_fast_gemul(const qb_gemul_plan &p, const GemulProg &g, int arith_bits);
const FastCGemul *find_fast_cgemul(const qb_cgemul_plan &p, const CGemulProg &g, int arith_bits);

}  // namespace qb
