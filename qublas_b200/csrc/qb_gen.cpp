// qb_gen — build-time generator: compiles the BASELINE plans and prints their compile-time kernel
// configurations (qb_codegen) into _build/qb_fast_gen.inc.
#include <cstdio>
#include "qb_baseline.h"
#include "qb_codegen.h"
using namespace qb;

int main()
{
    {
        qb_gemul_plan p = baseline_c3(); GemulProg g; Analysis an;
        if (compile_gemul(p, &g, an) || arith_bits(an) != 32) { fprintf(stderr, "c3: %s\n", an.err.c_str()); return 1; }
        sanitize32_stage(g.mul); sanitize32(g.tree); sanitize32_stage(g.out);
        printf("%s\n", emit_gemul_cfg("CfgC3", p, g).c_str());
    }
    {
        qb_cgemul_plan p = baseline_c4(); CGemulProg g; Analysis an;
        if (compile_cgemul(p, &g, an) || arith_bits(an) != 32) { fprintf(stderr, "c4: %s\n", an.err.c_str()); return 1; }
        sanitize32(g.tf); sanitize32(g.tree_re); sanitize32(g.tree_im); sanitize32_stage(g.out_re); sanitize32_stage(g.out_im);
        printf("%s\n", emit_cgemul_cfg("CfgC4", p, g).c_str());
    }
    for (int b = 0; b < 2; b++)
        for (int mode = 0; mode < 7; mode++) {
            qb_reduce_plan p = baseline_c2(b != 0, mode); ReduceProg r; Analysis an;
            if (compile_reduce(p, &r, an) || arith_bits(an) != 32) { fprintf(stderr, "c2: %s\n", an.err.c_str()); return 1; }
            sanitize32(r.tree);
            char name[32]; snprintf(name, sizeof(name), "CfgC2%c%d", b ? 'b' : 'a', mode);
            printf("%s\n", emit_reduce_cfg(name, p, r).c_str());
        }
    return 0;
}
