// qb_codegen.h — emits the C++ text of a compile-time kernel configuration from a compiled plan.
//
// This is the step that "lowers each Qu type's compile-time parameters into kernel template
// arguments": the POD programs of qb_prog.h are printed as constexpr initialisers inside a tag
// struct, and the kernel templates of qb_fast_*.cuh are instantiated on that tag.  Used at build
// time (qb_gen -> _build/qb_fast_gen.inc, the AOT-specialised BASELINE configurations).
#pragma once
#include <string>
#include "qb_plan.h"

namespace qb {

std::string emit_qstage(const QStage &q);
std::string emit_addstage(const AddStage &a);
// group = instances a CTA of the register-tile template works on at a time (1 unless the instance is small)
// epi: ANUS polynomial applied to every result element before it is stored (fused epilogue), or nullptr
std::string emit_gemul_cfg(const std::string &name, const qb_gemul_plan &p, const GemulProg &g, int group = 1, const PolyProg *epi = nullptr, const PolyProg *pro_a = nullptr, const PolyProg *pro_b = nullptr);
std::string emit_gemul_thread_cfg(const std::string &name, const qb_gemul_plan &p, const GemulProg &g, int arith_bits, const PolyProg *epi = nullptr, const PolyProg *pro_a = nullptr, const PolyProg *pro_b = nullptr);
std::string emit_polyprog(const PolyProg &P);
std::string emit_poly_cfg(const std::string &name, const PolyProg &prog, int xw, int ow, int arith_bits);
std::string emit_reduce_cfg(const std::string &name, const qb_reduce_plan &p, const ReduceProg &r);
std::string emit_reduce_any_cfg(const std::string &name, const qb_reduce_plan &p, const ReduceProg &r);
std::string emit_cgemul_cfg(const std::string &name, const qb_cgemul_plan &p, const CGemulProg &g);
std::string emit_cgemul_thread_cfg(const std::string &name, const qb_cgemul_plan &p, const CGemulProg &g, int arith_bits);
std::string emit_ew_cfg(const std::string &name, int kind, const TfProg *tf, const BcmProg *bcm, int op, const AddStage *stage, int xw, int yw, int ow,
                        int arith_bits);
std::string emit_plan_bytes(const void *plan, size_t n);

// The decisions to run a quantizer on the FP32 pipe (shared by the code generator and by qb200_selftest_fp, which sweeps
// every operand pair of exactly those nodes against the integer form).
bool gemul_fp_mul(const qb_gemul_plan &p, const GemulProg &g);
// dynamic shared memory of the register-tile template for `group` instances per CTA (0: does not fit); *sep: separate 32-bit copy
uint64_t gemul_tile_smem(const qb_gemul_plan &p, const GemulProg &g, int group, bool *sep);
void gemul_clamp_thresholds(const GemulProg &g, long long *p_hi, long long *p_lo);
struct CgemulFp { bool fa = false, fb = false, fc = false, fadd_ab = false, fadd_ba = false; Ival ab, cd, ba;
                  bool fbcm = false; };      // fbcm: all four BasicComplexMul products are exact as one FFMA each
CgemulFp cgemul_fp(const qb_cgemul_plan &p, const CGemulProg &g);

}  // namespace qb
