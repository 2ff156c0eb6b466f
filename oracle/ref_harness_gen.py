#!/usr/bin/env python3
"""Golden-vector generator: runs the UNMODIFIED reference header and records its answers.

TEST INFRASTRUCTURE.  Runs only in the build container (needs /root/reference, read-only):

    python oracle/ref_harness_gen.py            # regenerates tests/golden/ref_*.json

What it does
  1. builds a deterministic list of cases (seeded) over the hot path of the reference:
       scalar   Qmul / Qadd / Qsub / converting ctor       include/QuBLAS.h:3146-3235, 2398-2411
       reduce   Qreduce vector form and variadic form       include/QuBLAS.h:4903-5018
       tf       Qmul<TFComplexMul<...>>(x, y)               include/QuBLAS.h:3510-3535
       gemul    the readme.md:80-87 composition Qmul -> Qreduce -> converting assignment
       cgemul   TFComplexMul products -> variadic complex Qreduce (QuBLAS.h:3549-3568)
  2. emits oracle/_ref/ref_cases_*.cpp that instantiates every case with the reference's own
     templates, compiles it with g++ -std=c++23 against /root/reference/include (sources are
     compiled where they lie; nothing is copied), loads the .so and runs every case;
  3. writes inputs, the reference's outputs AND the reference's deduced result types to
     tests/golden/ref_*.json.  Those fixtures travel to the GPU box; the reference does not.

All cases stay inside the parity domain of SURVEY.md 2.3 (every intermediate <= 64 bits, no RND
right shift of 31/32/63/64, no unsigned WRP::TCPL target with intBits+fracBits in {32,64} — there
ArbiInt<32>::allOnes() sign-extends and the wrap mask is a no-op, found by this generator) where
the reference equals the ideal integer model.
"""
import ctypes
import json
import os
import random
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_INC = os.environ.get("QUBLAS_REF_INCLUDE", "/root/reference/include")
OUT_DIR = os.path.join(ROOT, "oracle", "_ref")
GOLD_DIR = os.path.join(ROOT, "tests", "golden")

QU = ["RND::POS_INF", "RND::NEG_INF", "RND::ZERO", "RND::INF", "RND::CONV", "TRN::TCPL", "TRN::SMGN"]
OF = ["SAT::TCPL", "SAT::ZERO", "SAT::SMGN", "WRP::TCPL"]


def T(i, f, s=True, q=5, o=0):
    return {"i": i, "f": f, "s": bool(s), "q": q, "o": o}


def cxx_type(t):
    return "Qu<intBits<%d>, fracBits<%d>, isSigned<%s>, QuMode<%s>, OfMode<%s>>" % (
        t["i"], t["f"], "true" if t["s"] else "false", QU[t["q"]], OF[t["o"]])


def cxx_tags(tags):
    """Partial tag list (merger arguments): any subset of i/f/s/q/o plus FullPrec."""
    if tags is None:
        return ""
    parts = []
    if "i" in tags: parts.append("intBits<%d>" % tags["i"])
    if "f" in tags: parts.append("fracBits<%d>" % tags["f"])
    if "s" in tags: parts.append("isSigned<%s>" % ("true" if tags["s"] else "false"))
    if "q" in tags: parts.append("QuMode<%s>" % QU[tags["q"]])
    if "o" in tags: parts.append("OfMode<%s>" % OF[tags["o"]])
    if tags.get("full"): parts.append("FullPrec")
    return ", ".join(parts)


def width(t):
    return 1 + t["i"] + t["f"]


def raw_range(t):
    w = width(t)
    return -(1 << (w - 1)), (1 << (w - 1)) - 1


def rand_raw(rng, t):
    lo, hi = raw_range(t)
    r = rng.random()
    if r < 0.08: return lo
    if r < 0.16: return hi
    if r < 0.22: return rng.choice([0, 1, -1, 2, -2])
    if r < 0.35: return rng.randint(max(lo, -40), min(hi, 40))
    return rng.randint(lo, hi)


BAD_SHIFTS = {31, 32, 63, 64}

# ------------------------------------------------------------------------------------------
# case construction
# ------------------------------------------------------------------------------------------

def scalar_cases(rng):
    cases = []
    for op in ("mul", "add", "sub"):
        for q in range(7):
            for o in range(4):
                for rep in range(2):
                    while True:
                        big = rep == 1 and rng.random() < 0.6
                        wmax = 30 if big else 15
                        a = T(rng.randint(-2, wmax // 2 + 3), 0, rng.random() < 0.8, rng.randrange(7), rng.randrange(4))
                        a["f"] = rng.randint(max(-a["i"], -3), wmax - a["i"])
                        b = T(rng.randint(-2, wmax // 2 + 3), 0, rng.random() < 0.8, rng.randrange(7), rng.randrange(4))
                        b["f"] = rng.randint(max(-b["i"], -3), wmax - b["i"])
                        if width(a) < 2 or width(b) < 2:
                            continue
                        from_f = a["f"] + b["f"] if op == "mul" else max(a["f"], b["f"])
                        s = rng.choice([-3, -1, 0, 1, 1, 2, 3, 5, 8, 12]) if rng.random() < 0.85 else rng.randint(13, 28)
                        to_f = from_f - s
                        if op == "mul":
                            inter = width(a) + width(b)
                        else:
                            mf = max(a["f"], b["f"])
                            inter = max(width(a) + mf - a["f"], width(b) + mf - b["f"]) + 1
                        inter += max(0, -s) + 1
                        if inter > 62 or s in BAD_SHIFTS:
                            continue
                        # result width: sometimes tight (forces overflow handling), sometimes roomy
                        full_int = inter - from_f
                        to_i = rng.choice([full_int - 6, full_int - 3, full_int - 1, full_int, full_int + 1])
                        to_i = max(to_i, -to_f)
                        if 1 + to_i + to_f < 2 or 1 + to_i + to_f > 62:
                            continue
                        to = T(to_i, to_f, rng.random() < 0.75, q, o)
                        if o == 3 and not to["s"] and (to_i + to_f) in (32, 64):
                            continue        # reference defect: unsigned WRP mask of exactly 32/64 bits is a no-op
                        break
                    cases.append({"kind": "scalar", "op": op, "a": a, "b": b, "to": to})
    # merger-default and partial-tag cases (QuBLAS.h:3107-3140)
    for op in ("mul", "add", "sub"):
        for tags in (None, {"full": True}, {"i": 9}, {"f": 2, "q": 4}, {"o": 3, "s": False}, {"full": True, "q": 0}):
            a = T(rng.randint(2, 7), rng.randint(0, 6), rng.random() < 0.8, rng.randrange(7), rng.randrange(4))
            b = dict(a) if rng.random() < 0.4 else T(rng.randint(2, 7), rng.randint(0, 6), rng.random() < 0.8,
                                                     rng.randrange(7), rng.randrange(4))
            cases.append({"kind": "scalar", "op": op, "a": a, "b": b, "tags": tags})
    # converting ctor
    for q in range(7):
        for o in range(4):
            for rep in range(2):
                while True:
                    a = T(rng.randint(0, 20), rng.randint(0, 20), True, rng.randrange(7), rng.randrange(4))
                    s = rng.choice([-4, -1, 0, 1, 2, 3, 4, 7, 9, 15])
                    to_f = a["f"] - s
                    to_i = rng.choice([a["i"] - 5, a["i"] - 2, a["i"], a["i"] + 2])
                    to_i = max(to_i, -to_f)
                    if 2 <= 1 + to_i + to_f <= 62 and width(a) + max(0, -s) <= 62 and s not in BAD_SHIFTS:
                        break
                cases.append({"kind": "scalar", "op": "conv", "a": a, "b": a, "to": T(to_i, to_f, rng.random() < 0.7, q, o)})
    return cases


def wide_cases(rng):
    """Two-word slice: Qmul / Qadd whose exact intermediate is a 65..126-bit ArbiInt in the
    reference.  `band` = the value compared by intConvert can reach [2^63, 2^64) (reference defect 4)."""
    cases = []
    for q in (0, 1, 2, 3, 5, 6):            # RND::CONV is defective on the wide path (defect 3)
        for o in range(4):
            for band in (False, True):
                while True:
                    wa, wb = rng.randint(33, 60), rng.randint(33, 60)
                    if not 66 <= wa + wb <= 120:
                        continue
                    fa, fb = rng.randint(0, wa - 2), rng.randint(0, wb - 2)
                    a = T(wa - 1 - fa, fa, True, rng.randrange(7), rng.randrange(4))
                    b = T(wb - 1 - fb, fb, True, rng.randrange(7), rng.randrange(4))
                    prod_bits = wa + wb
                    s = prod_bits - (rng.randint(40, 61) if not band else rng.randint(64, 70))
                    if s < 1 or s in BAD_SHIFTS or s > 62:
                        continue
                    to_f = fa + fb - s
                    to_w = rng.randint(24, 60)
                    to_i = to_w - 1 - to_f
                    if to_i + to_f < 1:
                        continue
                    to = T(to_i, to_f, rng.random() < 0.8, q, o)
                    if o == 3 and not to["s"] and (to_i + to_f) in (32, 64):
                        continue
                    break
                cases.append({"kind": "scalar", "op": "mul", "a": a, "b": b, "to": to, "wide": True, "band": band})
    # wide ADD: alignment shift pushes one operand beyond 64 bits
    for q in (0, 1, 2, 3, 5, 6):
        for o in range(4):
            while True:
                wa, wb = rng.randint(40, 60), rng.randint(20, 60)
                fa = rng.randint(0, 8); fb = fa + rng.randint(10, 30)            # a is shifted left by fb - fa
                if wa + (fb - fa) < 66 or wa + (fb - fa) > 120 or fb > wb - 2:
                    continue
                a = T(wa - 1 - fa, fa, True, rng.randrange(7), rng.randrange(4))
                b = T(wb - 1 - fb, fb, True, rng.randrange(7), rng.randrange(4))
                s = rng.randint(wa + fb - fa - 60, wa + fb - fa - 30)
                if s < 1 or s in BAD_SHIFTS or s > 62:
                    continue
                to_f = fb - s
                to_w = rng.randint(24, 60)
                to_i = to_w - 1 - to_f
                if to_i + to_f < 1:
                    continue
                to = T(to_i, to_f, True, q, o)
                break
            cases.append({"kind": "scalar", "op": "add", "a": a, "b": b, "to": to, "wide": True, "band": False})
    return cases


def level_list(rng, elem, nlev, uniform_mode=None):
    lv = []
    i, f = elem["i"], elem["f"]
    for l in range(nlev):
        i += rng.choice([0, 1, 1, 2])
        f -= rng.choice([0, 0, 1, 1, 2])
        if i + f < 1:
            f = 1 - i
        q = uniform_mode if uniform_mode is not None else rng.randrange(7)
        lv.append(T(i, f, rng.random() < 0.9, q, rng.randrange(4)))
    return lv


def reduce_cases(rng):
    cases = []
    vec_lens = [1, 2, 3, 4, 5, 6, 7, 8, 9, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 127, 128, 129, 255, 256, 257, 500, 512, 1000]
    for n in vec_lens:
        for rep in range(2):
            elem = T(rng.randint(3, 9), rng.randint(2, 9), rng.random() < 0.85, rng.randrange(7), rng.randrange(4))
            nlev = rng.choice([0, 1, 2, 3, 4, 6, 12]) if rep else rng.choice([1, 3, 5])
            cases.append({"kind": "reduce", "form": "vector", "len": n, "elem": elem,
                          "levels": level_list(rng, elem, nlev, None if rep else rng.randrange(7))})
    for n in [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 16, 17, 24, 31, 32, 33]:
        for rep in range(2):
            elem = T(rng.randint(3, 9), rng.randint(2, 9), True, rng.randrange(7), rng.randrange(4))
            nlev = rng.choice([0, 1, 2, 3, 6])
            cases.append({"kind": "reduce", "form": "variadic", "len": n, "elem": elem,
                          "levels": level_list(rng, elem, nlev)})
    # readme reductions (readme.md:57,61,64): red1 / red2 are pinned with their exact inputs below
    return cases


TF_TAGS = ["abT", "cdT", "baT", "abcT", "cdbT", "badT", "ABT", "BCT"]


def tf_cases(rng):
    cases = []
    t44 = T(4, 4)
    # the SURVEY 2.5 probe: distinct frac bits per tag expose the routing quirks
    cases.append({"kind": "tf", "x_re": t44, "x_im": t44, "y_re": t44, "y_im": t44,
                  "tags": {"abT": T(6, 3), "cdT": T(6, 2), "baT": T(6, 1), "abcT": T(10, 5), "cdbT": T(10, 6),
                           "badT": T(10, 7), "ABT": T(11, 3, True, 4, 0), "BCT": T(12, 2, True, 3, 0)}})
    cases.append({"kind": "tf", "x_re": t44, "x_im": t44, "y_re": t44, "y_im": t44, "tags": {}})
    # BASELINE config 4 types
    t = T(3, 12); P = T(6, 12); S = T(7, 12)
    cases.append({"kind": "tf", "x_re": t, "x_im": t, "y_re": t, "y_im": t,
                  "tags": {"abT": T(4, 12), "cdT": T(4, 11), "abcT": P, "cdbT": P, "badT": P, "ABT": S, "BCT": S}})
    # readme.md:92-109: Qcomplex<type1,type2> operands with the readme's TFComplexMul tags
    type1 = T(6, 3, True, 5, 1); type2 = T(6, -3)
    cases.append({"kind": "tf", "x_re": type1, "x_im": type2, "y_re": type1, "y_im": type2, "readme": True,
                  "tags": {"abT": type2, "cdT": type1, "baT": type2, "badT": type2, "ABT": type2, "BCT": type2}})
    for rep in range(12):
        def rt(lo=2, hi=8):
            return T(rng.randint(lo, hi), rng.randint(0, 8), rng.random() < 0.85, rng.randrange(7), rng.randrange(4))
        xr, xi, yr, yi = rt(), rt(), rt(), rt()
        if rep % 3 == 0:
            xi = dict(xr); yi = dict(yr)
        tags = {}
        for name in TF_TAGS:
            if rng.random() < 0.7:
                tags[name] = rt(4, 14)
        cases.append({"kind": "tf", "x_re": xr, "x_im": xi, "y_re": yr, "y_im": yi, "tags": tags})
    return cases


BCM_TAGS = ["acT", "bdT", "adT", "bcT", "acbdT", "adbcT"]


def bcm_cases(rng):
    """Qmul<BasicComplexMul<...>>(x, y) and the default Qmul(x, y) on complex scalars (QuBLAS.h:3422-3446)."""
    cases = []
    t44 = T(4, 4)
    cases.append({"kind": "bcm", "x_re": t44, "x_im": t44, "y_re": t44, "y_im": t44, "tags": {}, "default": True})
    for rep in range(10):
        def rt(lo=2, hi=8):
            return T(rng.randint(lo, hi), rng.randint(0, 8), rng.random() < 0.85, rng.randrange(7), rng.randrange(4))
        xr, xi, yr, yi = rt(), rt(), rt(), rt()
        tags = {}
        for name in BCM_TAGS:
            if rng.random() < 0.7:
                tags[name] = rt(4, 14)
        if len(tags) == 1:      # a pack of exactly ONE tag does not compile upstream (tagExtractor's single-argument
            tags = {}           # specialisation strips the wrapper, QuBLAS.h:158-162,173-177), so it cannot occur in user code
        cases.append({"kind": "bcm", "x_re": xr, "x_im": xi, "y_re": yr, "y_im": yi, "tags": tags, "default": False})
    return cases


def _rc_targs(rng, rt, allow_split):
    mode = rng.choice(["none", "one", "pair", "tags"] if allow_split else ["none", "one"])
    if mode == "none":
        return {"mode": "none"}
    if mode == "one":
        return {"mode": "one", "t": rt(4, 14)}
    return {"mode": mode, "re": rt(4, 14), "im": rt(4, 14)}


def rc_cases(rng):
    """real (x) complex scalar Qmul / Qadd / Qsub in both operand orders (QuBLAS.h:3600-3696), complex +/- complex with
    realT / imagT / positional targets (:3549-3590), and dot products of real-by-complex Qmul accumulated by the
    variadic complex Qreduce — the composition a mixed real/complex Qgemul performs."""
    def rt(lo=2, hi=8):
        return T(rng.randint(lo, hi), rng.randint(0, 8), rng.random() < 0.85, rng.randrange(7), rng.randrange(4))
    cases = []
    for op in ("mul", "add", "sub"):
        for real_first in (True, False):
            for rep in range(5):
                cases.append({"kind": "rc", "op": op, "real_first": real_first, "r": rt(), "c_re": rt(), "c_im": rt(),
                              "targs": _rc_targs(rng, rt, op == "mul")})
    for op in ("cadd", "csub"):
        for rep in range(6):
            cases.append({"kind": "rc", "op": op, "x_re": rt(), "x_im": rt(), "y_re": rt(), "y_im": rt(), "targs": _rc_targs(rng, rt, True)})
    for (K, real_first) in [(4, True), (8, False), (5, True), (7, False), (16, True), (3, False)]:
        r, cre, cim = T(rng.randint(3, 6), rng.randint(4, 8)), T(rng.randint(3, 6), rng.randint(4, 8)), T(rng.randint(3, 6), rng.randint(4, 8))
        targs = _rc_targs(rng, lambda lo, hi: T(rng.randint(7, 10), rng.randint(5, 9), True, rng.randrange(7), rng.randrange(4)), True)
        lv = [T(9 + l, 7 - l, True, rng.randrange(7), rng.randrange(3)) for l in range(rng.randint(1, 4))]
        cases.append({"kind": "rcgemul", "K": K, "real_first": real_first, "r": r, "c_re": cre, "c_im": cim, "targs": targs, "levels": lv})
    return cases


def gemul_cases(rng):
    cases = []
    type1 = T(6, 3, True, 5, 1)     # readme.md:31  Qu<isSigned<true>,intBits<6>,fracBits<3>,OfMode<SAT::ZERO>>
    type2 = T(6, -3, True, 5, 0)    # readme.md:35  Qu<intBits<6>,fracBits<-3>>
    for ta in (True, False):
        cases.append({"kind": "gemul", "M": 4, "N": 4, "K": 4, "a": type1, "b": type1, "mul": type1, "c": type1,
                      "levels": [type1, type2], "trans_a": ta, "form": "vector", "readme": True})
    for (M, N, K, form) in [(3, 2, 8, "vector"), (2, 3, 5, "vector"), (2, 3, 5, "variadic"), (5, 4, 16, "vector"),
                            (2, 2, 7, "variadic"), (2, 2, 7, "vector"), (3, 3, 1, "vector"), (4, 2, 32, "vector")]:
        a = T(rng.randint(3, 7), rng.randint(3, 8), True, rng.randrange(7), rng.randrange(4))
        b = T(rng.randint(3, 7), rng.randint(3, 8), True, rng.randrange(7), rng.randrange(4))
        mul = T(a["i"] + b["i"] - rng.randint(0, 2), max(a["f"], b["f"]) - rng.randint(0, 2), True, rng.randrange(7), rng.randrange(4))
        lv = level_list(rng, mul, rng.choice([1, 2, 3, 5]))
        c = T(mul["i"] + rng.randint(0, 3), mul["f"] - rng.randint(0, 4), True, rng.randrange(7), rng.randrange(4))
        cases.append({"kind": "gemul", "M": M, "N": N, "K": K, "a": a, "b": b, "mul": mul, "c": c, "levels": lv,
                      "trans_a": rng.random() < 0.5, "form": form})
    return cases


def fromdouble_cases(rng):
    """Qu_s(double) (QuBLAS.h:2387-2393).  RND::CONV is excluded: the reference rounds the 2400-bit load
    buffer through its wide path, where CONV is defective."""
    cases = []
    for q in (0, 1, 2, 3, 5, 6):
        for o in range(4):
            i, f = rng.randint(2, 12), rng.randint(0, 14)
            cases.append({"kind": "fromdouble", "to": T(i, f, rng.random() < 0.8, q, o)})
    cases.append({"kind": "fromdouble", "to": T(6, -3)})        # readme.md:35  type2 q2 = 16
    return cases


def fromdouble64_cases(rng):
    """Round 2: Qu_s(double) into 64-bit-wide targets (intBits + fracBits = 63) and neighbours, with |v| around 2^62..2^64
    in units of the target's lsb -- the band where a 64-bit implementation must decide the overflow on the exact value."""
    cases = []
    for (i, f) in ((63, 0), (40, 23), (62, 0), (50, 13)):
        for q in (0, 1, 2, 3, 5, 6):
            for o in range(4):
                cases.append({"kind": "fromdouble", "to": T(i, f, True, q, o), "big": True})
    return cases


def cgemul_cases(rng):
    cases = []
    for (K, half) in [(4, False), (8, True), (16, False), (5, False)]:
        t = T(3, 12); P = T(6, 12); S = T(7, 12)
        tags = {"abT": T(4, 12), "cdT": T(4, 11), "abcT": P, "cdbT": P, "badT": P, "ABT": S, "BCT": S}
        lv = [T(8 + l, 12 - l, True, 4, 0) for l in range(4)]
        cases.append({"kind": "cgemul", "K": K, "x_re": t, "x_im": t, "y_re": t, "y_im": t, "tags": tags,
                      "levels": lv, "half_range": half})
    return cases


# ------------------------------------------------------------------------------------------
# C++ emission
# ------------------------------------------------------------------------------------------

PRELUDE = r'''// GENERATED by oracle/ref_harness_gen.py — do not commit. Includes the unmodified reference header.
#include "QuBLAS.h"
#include <cstdint>
using namespace QuBLAS;
template <typename Q> static inline void put_type(int64_t *o, const Q &) {
    o[0] = Q::intB; o[1] = Q::fracB; o[2] = Q::isS; o[3] = Q::QuM; o[4] = Q::OfM;
}
template <typename Q> static inline Q mk(int64_t raw) { Q q; q.data.data = (decltype(q.data.data))raw; return q; }
'''


def emit_scalar(c, idx):
    A, B = cxx_type(c["a"]), cxx_type(c["b"])
    if c["op"] == "conv":
        body = "auto r = %s(mk<%s>(in[0]));" % (cxx_type(c["to"]), A)
    else:
        fn = {"mul": "Qmul", "add": "Qadd", "sub": "Qsub"}[c["op"]]
        targ = cxx_type(c["to"]) if "to" in c else cxx_tags(c.get("tags"))
        body = "auto r = %s<%s>(mk<%s>(in[0]), mk<%s>(in[1]));" % (fn, targ, A, B)
    return "static void case_%d(const int64_t *in, int64_t *out) { %s out[0] = (int64_t)r.data.data; put_type(out + 1, r); }\n" % (idx, body)


def emit_levels(levels):
    if not levels:
        return ""
    return "TypeList<%s>" % ", ".join(cxx_type(t) for t in levels)


def emit_reduce(c, idx):
    E = cxx_type(c["elem"])
    n = c["len"]
    L = emit_levels(c["levels"])
    if c["form"] == "vector":
        body = ("static Qu<dim<%d>, %s> v; for (int i = 0; i < %d; i++) v[i].data.data = (decltype(v[i].data.data))in[i]; "
                "auto r = Qreduce<%s>(v);" % (n, E, n, L))
    else:
        args = ", ".join("mk<%s>(in[%d])" % (E, i) for i in range(n))
        body = "auto r = Qreduce<%s>(%s);" % (L, args)
    return "static void case_%d(const int64_t *in, int64_t *out) { %s out[0] = (int64_t)r.data.data; put_type(out + 1, r); }\n" % (idx, body)


def emit_tf_args(tags):
    return ", ".join("%s<%s>" % (k, cxx_type(v)) for k, v in tags.items())


def emit_tf(c, idx):
    X = "Qcomplex<%s, %s>" % (cxx_type(c["x_re"]), cxx_type(c["x_im"]))
    Y = "Qcomplex<%s, %s>" % (cxx_type(c["y_re"]), cxx_type(c["y_im"]))
    body = ("%s x; x.real = mk<%s>(in[0]); x.imag = mk<%s>(in[1]); %s y; y.real = mk<%s>(in[2]); y.imag = mk<%s>(in[3]); "
            "auto r = Qmul<TFComplexMul<%s>>(x, y); out[0] = (int64_t)r.real.data.data; out[1] = (int64_t)r.imag.data.data; "
            "put_type(out + 2, r.real); put_type(out + 7, r.imag);"
            % (X, cxx_type(c["x_re"]), cxx_type(c["x_im"]), Y, cxx_type(c["y_re"]), cxx_type(c["y_im"]), emit_tf_args(c["tags"])))
    return "static void case_%d(const int64_t *in, int64_t *out) { %s }\n" % (idx, body)


def emit_bcm(c, idx):
    X = "Qcomplex<%s, %s>" % (cxx_type(c["x_re"]), cxx_type(c["x_im"]))
    Y = "Qcomplex<%s, %s>" % (cxx_type(c["y_re"]), cxx_type(c["y_im"]))
    call = "Qmul(x, y)" if c["default"] else "Qmul<BasicComplexMul<%s>>(x, y)" % emit_tf_args(c["tags"])
    body = ("%s x; x.real = mk<%s>(in[0]); x.imag = mk<%s>(in[1]); %s y; y.real = mk<%s>(in[2]); y.imag = mk<%s>(in[3]); "
            "auto r = %s; out[0] = (int64_t)r.real.data.data; out[1] = (int64_t)r.imag.data.data; "
            "put_type(out + 2, r.real); put_type(out + 7, r.imag);"
            % (X, cxx_type(c["x_re"]), cxx_type(c["x_im"]), Y, cxx_type(c["y_re"]), cxx_type(c["y_im"]), call))
    return "static void case_%d(const int64_t *in, int64_t *out) { %s }\n" % (idx, body)


def emit_gemul(c, idx):
    M, N, K = c["M"], c["N"], c["K"]
    A, B, P, C = (cxx_type(c[k]) for k in ("a", "b", "mul", "c"))
    L = emit_levels(c["levels"])
    lines = ["static void case_%d(const int64_t *in, int64_t *out) {" % idx,
             "  const int64_t *Ain = in, *Bin = in + %d;" % (M * K),
             "  for (int j = 0; j < %d; j++) for (int i = 0; i < %d; i++) {" % (N, M)]
    a_idx = "k + i * %d" % K if c["trans_a"] else "i + k * %d" % M
    if c["form"] == "vector":
        lines.append("    static Qu<dim<%d>, %s> p;" % (K, P))
        lines.append("    for (int k = 0; k < %d; k++) p[k] = Qmul<%s>(mk<%s>(Ain[%s]), mk<%s>(Bin[k + j * %d]));" % (K, P, A, a_idx, B, K))
        lines.append("    %s cij = Qreduce<%s>(p);" % (C, L))
    else:
        args = ", ".join("Qmul<%s>(mk<%s>(Ain[%s]), mk<%s>(Bin[%d + j * %d]))" % (
            P, A, a_idx.replace("k", str(k)), B, k, K) for k in range(K))
        lines.append("    %s cij = Qreduce<%s>(%s);" % (C, L, args))
    lines.append("    out[i + j * %d] = (int64_t)cij.data.data;" % M)
    lines.append("  }\n}\n")
    return "\n".join(lines)


def emit_cgemul(c, idx):
    """One output of a complex row dot product: variadic complex Qreduce over TF products."""
    K = c["K"]
    X = "Qcomplex<%s, %s>" % (cxx_type(c["x_re"]), cxx_type(c["x_im"]))
    Y = "Qcomplex<%s, %s>" % (cxx_type(c["y_re"]), cxx_type(c["y_im"]))
    L = emit_levels(c["levels"])
    lines = ["static void case_%d(const int64_t *in, int64_t *out) {" % idx,
             "  %s x[%d]; %s y[%d];" % (X, K, Y, K),
             "  for (int k = 0; k < %d; k++) { x[k].real = mk<%s>(in[4*k]); x[k].imag = mk<%s>(in[4*k+1]); "
             "y[k].real = mk<%s>(in[4*k+2]); y[k].imag = mk<%s>(in[4*k+3]); }" % (
                 K, cxx_type(c["x_re"]), cxx_type(c["x_im"]), cxx_type(c["y_re"]), cxx_type(c["y_im"]))]
    args = ", ".join("Qmul<TFComplexMul<%s>>(x[%d], y[%d])" % (emit_tf_args(c["tags"]), k, k) for k in range(K))
    lines.append("  auto r = Qreduce<%s>(%s);" % (L, args))
    lines.append("  out[0] = (int64_t)r.real.data.data; out[1] = (int64_t)r.imag.data.data; put_type(out + 2, r.real); put_type(out + 7, r.imag);")
    lines.append("}\n")
    return "\n".join(lines)


def emit_targs(t):
    if t["mode"] == "none": return ""
    if t["mode"] == "one": return cxx_type(t["t"])
    if t["mode"] == "pair": return "%s, %s" % (cxx_type(t["re"]), cxx_type(t["im"]))
    return "realT<%s>, imagT<%s>" % (cxx_type(t["re"]), cxx_type(t["im"]))


_PUT_C = ("out[0] = (int64_t)r.real.data.data; out[1] = (int64_t)r.imag.data.data; put_type(out + 2, r.real); put_type(out + 7, r.imag);")


def emit_rc(c, idx):
    ta = emit_targs(c["targs"])
    if c["op"] in ("cadd", "csub"):
        X = "Qcomplex<%s, %s>" % (cxx_type(c["x_re"]), cxx_type(c["x_im"]))
        Y = "Qcomplex<%s, %s>" % (cxx_type(c["y_re"]), cxx_type(c["y_im"]))
        body = ("%s x; x.real = mk<%s>(in[0]); x.imag = mk<%s>(in[1]); %s y; y.real = mk<%s>(in[2]); y.imag = mk<%s>(in[3]); auto r = %s<%s>(x, y); %s"
                % (X, cxx_type(c["x_re"]), cxx_type(c["x_im"]), Y, cxx_type(c["y_re"]), cxx_type(c["y_im"]),
                   "Qadd" if c["op"] == "cadd" else "Qsub", ta, _PUT_C))
    else:
        Cx = "Qcomplex<%s, %s>" % (cxx_type(c["c_re"]), cxx_type(c["c_im"]))
        fn = {"mul": "Qmul", "add": "Qadd", "sub": "Qsub"}[c["op"]]
        args = "s, z" if c["real_first"] else "z, s"
        body = ("auto s = mk<%s>(in[0]); %s z; z.real = mk<%s>(in[1]); z.imag = mk<%s>(in[2]); auto r = %s<%s>(%s); %s"
                % (cxx_type(c["r"]), Cx, cxx_type(c["c_re"]), cxx_type(c["c_im"]), fn, ta, args, _PUT_C))
    return "static void case_%d(const int64_t *in, int64_t *out) { %s }\n" % (idx, body)


def emit_rcgemul(c, idx):
    """One output of a mixed real/complex dot product: variadic complex Qreduce over real-by-complex Qmul."""
    K = c["K"]
    Cx = "Qcomplex<%s, %s>" % (cxx_type(c["c_re"]), cxx_type(c["c_im"]))
    ta = emit_targs(c["targs"])
    lines = ["static void case_%d(const int64_t *in, int64_t *out) {" % idx,
             "  %s s[%d]; %s z[%d];" % (cxx_type(c["r"]), K, Cx, K),
             "  for (int k = 0; k < %d; k++) { s[k] = mk<%s>(in[3*k]); z[k].real = mk<%s>(in[3*k+1]); z[k].imag = mk<%s>(in[3*k+2]); }"
             % (K, cxx_type(c["r"]), cxx_type(c["c_re"]), cxx_type(c["c_im"]))]
    args = ", ".join(("Qmul<%s>(s[%d], z[%d])" if c["real_first"] else "Qmul<%s>(z[%d], s[%d])") % (ta, k, k) for k in range(K))
    lines.append("  auto r = Qreduce<%s>(%s);" % (emit_levels(c["levels"]), args))
    lines.append("  " + _PUT_C)
    lines.append("}\n")
    return "\n".join(lines)


def emit_fromdouble(c, idx):
    body = ("double d; __builtin_memcpy(&d, &in[0], 8); %s r = d;" % cxx_type(c["to"]))
    return "static void case_%d(const int64_t *in, int64_t *out) { %s out[0] = (int64_t)r.data.data; put_type(out + 1, r); }\n" % (idx, body)


EMIT = {"bcm": emit_bcm, "fromdouble": emit_fromdouble, "scalar": emit_scalar, "reduce": emit_reduce, "tf": emit_tf, "gemul": emit_gemul, "cgemul": emit_cgemul,
        "rc": emit_rc, "rcgemul": emit_rcgemul}


def n_inputs(c):
    k = c["kind"]
    if k == "scalar": return 2
    if k == "fromdouble": return 1
    if k == "reduce": return c["len"]
    if k == "tf" or k == "bcm": return 4
    if k == "gemul": return c["M"] * c["K"] + c["K"] * c["N"]
    if k == "cgemul": return 4 * c["K"]
    if k == "rc": return 4 if c["op"] in ("cadd", "csub") else 3
    if k == "rcgemul": return 3 * c["K"]


def n_outputs(c):
    k = c["kind"]
    if k == "scalar" or k == "reduce" or k == "fromdouble": return 6
    if k in ("tf", "cgemul", "bcm", "rc", "rcgemul"): return 12
    if k == "gemul": return c["M"] * c["N"]


def _dbits(x):
    import struct
    return struct.unpack("<q", struct.pack("<d", float(x)))[0]


def gen_inputs(rng, c):
    k = c["kind"]
    if k == "fromdouble" and c.get("big"):
        t = c["to"]
        ulp = 2.0 ** (-t["f"])
        W = t["i"] + t["f"]                       # value bits; max raw = 2^W - 1
        mags = [2.0 ** (W - 1), 2.0 ** (W - 1) + 2.0 ** (W - 40), 2.0 ** W - 2.0 ** (W - 52), 2.0 ** W, 2.0 ** W + 2.0 ** (W - 50),
                1.5 * 2.0 ** W, 2.0 ** (W + 1) - 2.0 ** (W - 51), 2.0 ** (W + 1), 3.0 * 2.0 ** (W + 1), 2.0 ** (W + 7), 2.0 ** (W + 30), 1e30, 1e300]
        mags += [rng.uniform(2.0 ** (W - 2), 2.0 ** (W + 2)) for _ in range(10)]
        vals = []
        for m in mags:
            vals += [m * ulp, -m * ulp]
        return [[_dbits(v)] for v in vals]
    if k == "fromdouble":
        t = c["to"]
        ulp = 2.0 ** (-t["f"])
        top = 2.0 ** t["i"]
        vals = [0.0, 1.0, -1.0, 1.25, -1.25, 16.0, 0.5 * ulp, -0.5 * ulp, 1.5 * ulp, -1.5 * ulp, 2.5 * ulp, -2.5 * ulp, 0.49999 * ulp,
                -0.50001 * ulp, top, -top, top - ulp, -top - ulp, top * 3.7, -top * 5.1, 1e-9, -1e-9, 3.141592653589793, -2.718281828459045]
        vals += [rng.uniform(-1.3 * top, 1.3 * top) for _ in range(20)]
        vals += [(rng.randint(-2 ** 10, 2 ** 10) + 0.5) * ulp for _ in range(8)]        # exact ties
        return [[_dbits(v)] for v in vals]
    if k == "scalar":
        return [[rand_raw(rng, c["a"]), rand_raw(rng, c["b"])] for _ in range(48)]
    if k == "reduce":
        reps = 6 if c["len"] <= 64 else 3
        return [[rand_raw(rng, c["elem"]) for _ in range(c["len"])] for _ in range(reps)]
    if k == "bcm":
        return [[rand_raw(rng, c["x_re"]), rand_raw(rng, c["x_im"]), rand_raw(rng, c["y_re"]), rand_raw(rng, c["y_im"])] for _ in range(48)]
    if k == "tf":
        sets = [[rand_raw(rng, c["x_re"]), rand_raw(rng, c["x_im"]), rand_raw(rng, c["y_re"]), rand_raw(rng, c["y_im"])]
                for _ in range(48)]
        if c["x_re"] == T(4, 4) and c["tags"].get("abT") == T(6, 3):
            sets.insert(0, [37, -22, -51, 19])          # SURVEY 2.5 golden
        if c.get("readme"):
            sets.insert(0, [8, 0, 8, 0])                # complex_vec = {a, a, a}, a = 1.0
            sets.insert(0, [8, 0, 24, 0])               # complex1 = {1, 2}, complex2 = {3, 4}
        return sets
    if k == "gemul":
        if c.get("readme"):
            m1 = [8 * v for v in range(1, 17)]          # readme.md:43-46, raw = value * 2^3, column-major
            return [m1 + m1]
        return [[rand_raw(rng, c["a"]) for _ in range(c["M"] * c["K"])] + [rand_raw(rng, c["b"]) for _ in range(c["K"] * c["N"])]
                for _ in range(4)]
    if k == "rc":
        ts = [c[n] for n in (("x_re", "x_im", "y_re", "y_im") if c["op"] in ("cadd", "csub") else ("r", "c_re", "c_im"))]
        return [[rand_raw(rng, t) for t in ts] for _ in range(48)]
    if k == "rcgemul":
        return [[rand_raw(rng, c[n]) for _ in range(c["K"]) for n in ("r", "c_re", "c_im")] for _ in range(8)]
    if k == "cgemul":
        out = []
        for _ in range(8):
            row = []
            for kk in range(c["K"]):
                for t in (c["x_re"], c["x_im"], c["y_re"], c["y_im"]):
                    v = rand_raw(rng, t)
                    if c["half_range"]:
                        v //= 2
                    row.append(v)
            out.append(row)
        return out


def build_harness(cases, include_dir, out_dir, tag, n_tu=8, extra_flags=()):
    """Emit + compile the case table against the QuBLAS.h found in include_dir; returns run(idx, inputs)."""
    os.makedirs(out_dir, exist_ok=True)
    tus = [[] for _ in range(n_tu)]
    for idx, c in enumerate(cases):
        tus[idx % n_tu].append(idx)
    objs, procs = [], []
    for t, idxs in enumerate(tus):
        src = os.path.join(out_dir, "%s_cases_%d.cpp" % (tag, t))
        with open(src, "w") as f:
            f.write(PRELUDE)
            for idx in idxs:
                f.write(EMIT[cases[idx]["kind"]](cases[idx], idx))
            f.write('extern "C" int ref_run_%d(int id, const int64_t *in, int64_t *out) {\n  switch (id) {\n' % t)
            for idx in idxs:
                f.write("  case %d: case_%d(in, out); return 0;\n" % (idx, idx))
            f.write("  default: return -1; }\n}\n")
        obj = src[:-4] + ".o"
        objs.append(obj)
        procs.append(subprocess.Popen(["g++", "-std=c++23", "-O1", "-fPIC", "-w", "-I", include_dir, "-c", src, "-o", obj]))
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("harness failed to compile against " + include_dir)
    so = os.path.join(out_dir, "libq%s_cases.so" % tag)
    subprocess.check_call(["g++", "-shared", "-o", so] + objs + list(extra_flags))
    lib = ctypes.CDLL(so)

    def run(idx, row):
        fn = getattr(lib, "ref_run_%d" % (idx % n_tu))
        fn.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]
        a = (ctypes.c_int64 * len(row))(*row)
        o = (ctypes.c_int64 * n_outputs(cases[idx]))()
        rc = fn(idx, a, o)
        if rc != 0:
            raise RuntimeError("case %d failed" % idx)
        return list(o)
    return run


def all_cases(rng):
    return scalar_cases(rng) + reduce_cases(rng) + tf_cases(rng) + gemul_cases(rng) + cgemul_cases(rng) + fromdouble_cases(rng) + bcm_cases(rng) + wide_cases(rng)


TEXT_GOLDENS = {"bitstream_example": "ref_bitstream.txt", "anus_example": "ref_anus.txt"}


def text_goldens():
    """tests/cpp/<name>.cpp compiled against the REFERENCE header; its stdout is the fixture."""
    os.makedirs(OUT_DIR, exist_ok=True)
    for name, gold in TEXT_GOLDENS.items():
        src = os.path.join(os.path.dirname(GOLD_DIR), "cpp", name + ".cpp")
        exe = os.path.join(OUT_DIR, name + "_ref")
        subprocess.check_call(["g++", "-std=c++23", "-O0", "-w", "-I" + REF_INC, src, "-o", exe])
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
        path = os.path.join(GOLD_DIR, gold)
        with open(path, "w") as f:
            f.write(out)
        print("wrote", path, len(out.splitlines()), "lines")


def main():
    if "--text" in sys.argv[1:]:
        return text_goldens()
    if "--fd64" in sys.argv[1:]:            # round 2: 64-bit-wide Qu_s(double) targets; own seed, own fixture
        rng = random.Random(20261020)
        cases = fromdouble64_cases(rng)
    elif "--rc" in sys.argv[1:]:            # added later: own seed, own fixture, the earlier fixtures stay byte-identical
        rng = random.Random(20261019)
        cases = rc_cases(rng)
    else:
        rng = random.Random(20261018)
        cases = all_cases(rng)
    os.makedirs(GOLD_DIR, exist_ok=True)
    run = build_harness(cases, REF_INC, OUT_DIR, "ref")
    fixtures = {}
    for idx, c in enumerate(cases):
        ins = gen_inputs(rng, c)
        outs = []
        for row in ins:
            assert len(row) == n_inputs(c)
            outs.append(run(idx, row))
        rec = dict(c)
        rec["inputs"] = ins
        rec["outputs"] = outs
        fixtures.setdefault("wide" if c.get("wide") else ("fromdouble64" if c.get("big") else ("rc" if c["kind"] == "rcgemul" else c["kind"])), []).append(rec)
    for kind, recs in fixtures.items():
        path = os.path.join(GOLD_DIR, "ref_%s.json" % kind)
        with open(path, "w") as f:
            json.dump({"generator": "oracle/ref_harness_gen.py", "reference": "pikapuma/QuBLAS include/QuBLAS.h (g++ 13.3 -std=c++23)",
                       "cases": recs}, f, separators=(",", ":"))
        print("wrote", path, len(recs), "cases")


if __name__ == "__main__":
    main()
