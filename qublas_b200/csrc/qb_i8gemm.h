// qb_i8gemm.h — tcgen05 kind::i8 tensor-core path (qb_i8gemm.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "qb_prog.h"

namespace qb {
bool i8gemm_shape_ok(uint32_t M, uint32_t N, uint32_t K);
void i8gemm_padded_dims(uint32_t M, uint32_t N, uint32_t K, uint32_t *Mp, uint32_t *Np, uint32_t *Kp);   // staging: Mp x Kp and Np x Kp bytes
bool i8gemm_zero_copy(const void *ptr, int word_bytes, bool k_major, uint32_t K);   // operand read in place, no pack launch
// words of the flag buffer of the fused operand narrowing (k_i8gemm_pair_fused); 0 = separate pack launches
size_t i8gemm_sync_words(const void *A, const void *B, uint32_t M, uint32_t N, uint32_t K, bool trans_a, bool trans_b, int a_wb, int b_wb);
// Fan-out of the result (GEMM -> all-gather in one kernel): the epilogue stores every element into the column-major matrices
// out[0 .. n) (leading dimension ldc, this call's rows starting at row0) -- peer memory over NVLink for the other devices.
struct I8FanOut { uint32_t n, row0; uint64_t ldc; void *out[8]; };
int  i8gemm_run(const void *A, const void *B, void *C, uint32_t M, uint32_t N, uint32_t K, bool trans_a, bool trans_b, int a_wb,
                int b_wb, int c_wb, const QStage &outq, int8_t *Ap, int8_t *Bp, uint32_t *sync, cudaStream_t s, const char **err, const I8FanOut *fan = nullptr);
}
