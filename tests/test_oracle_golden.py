"""Pins the CPU oracle (oracle/qb_oracle.c) and the host type resolver (qublas_b200/qtypes.py)
against (1) the reference's own 40 known-answer tests and (2) golden vectors produced by the
unmodified reference header (oracle/ref_harness_gen.py).  CPU only."""
import numpy as np
import pytest

from oracle import oracle
from qublas_b200 import abi
from qublas_b200.qtypes import (QType, TREE_VARIADIC, TREE_VECTOR, add_merge, mul_merge, resolve_bcm, resolve_tf,
                                tree_out_type)
from tests.helpers import (cgemul_plan, form_id, gemul_plan, load_cases, np_words, qt, reduce_plan, tf_tags,
                           tfcmul_plan)


def type_tuple(t: QType):
    return [t.i, t.f, int(t.s), t.q, t.o]


def test_reference_known_answer_tests():
    """test/RND/*.cpp, test/TRN/*.cpp of the reference: +-1.25 and the CONV ties."""
    cases = load_cases("ref_kat.json")
    assert len(cases) == 40
    for c in cases:
        got = oracle.conv128(int(c["raw"]), c["from"]["f"], qt(c["to"]))
        assert got == c["expect_raw"], c
        assert got / 2 ** c["to"]["f"] == c["expect_value"]


def test_scalar_ops_match_reference():
    cases = load_cases("ref_scalar.json")
    n = 0
    for c in cases:
        a, b = qt(c["a"]), qt(c["b"])
        if c["op"] == "conv":
            to = qt(c["to"])
        elif "to" in c:
            to = qt(c["to"])
        else:
            to = (mul_merge if c["op"] == "mul" else add_merge)(a, b, c.get("tags"))
        for inp, out in zip(c["inputs"], c["outputs"]):
            assert out[1:6] == type_tuple(to), (c, out)          # resolver == reference's deduced type
            if c["op"] == "conv":
                got = oracle.conv(inp[0], a.f, to)
            else:
                got = {"mul": oracle.qmul, "add": oracle.qadd, "sub": oracle.qsub}[c["op"]](inp[0], a, inp[1], b, to)
            assert got == out[0], (c["op"], c["a"], c["b"], to, inp, out[0], got)
            n += 1
    assert n > 10000


def test_wide_two_word_slice_matches_reference_where_fenced_in():
    """65..126-bit intermediates: the ideal model equals the reference for non-CONV modes as long as the value at
    the overflow compare stays below 2^63; the `band` cases document where the reference itself is defective."""
    cases = load_cases("ref_wide.json")
    safe = band_bad = 0
    for c in cases:
        a, b, to = qt(c["a"]), qt(c["b"]), qt(c["to"])
        f = {"mul": oracle.qmul, "add": oracle.qadd}[c["op"]]
        bad = sum(1 for inp, out in zip(c["inputs"], c["outputs"]) if f(inp[0], a, inp[1], b, to) != out[0])
        if c["band"]:
            band_bad += bad > 0
        else:
            assert bad == 0, (c["op"], c["a"], c["b"], c["to"])
            safe += 1
    assert safe == 48 and band_bad >= 6          # defect 4 is real: about half of the band cases differ


def test_all_modes_covered():
    cases = [c for c in load_cases("ref_scalar.json") if "to" in c]
    seen = {(c["op"], c["to"]["q"], c["to"]["o"]) for c in cases}
    for op in ("mul", "add", "sub", "conv"):
        for q in range(7):
            for o in range(4):
                assert (op, q, o) in seen


def test_reduce_matches_reference():
    cases = load_cases("ref_reduce.json")
    forms = set()
    for c in cases:
        elem = qt(c["elem"]); lv = [qt(t) for t in c["levels"]]
        form = form_id(c["form"]); forms.add((c["form"], c["len"] & 1))
        plan = reduce_plan(elem, lv, c["len"], form)
        ot = tree_out_type(c["len"], elem, lv, form)
        for inp, out in zip(c["inputs"], c["outputs"]):
            assert out[1:6] == type_tuple(ot), (c["form"], c["len"], out, ot)
            got, _ = oracle.reduce_batched(plan, np_words(inp, elem), 1)
            assert int(got[0]) == out[0], (c["form"], c["len"], c["elem"], c["levels"], out[0], int(got[0]))
    assert forms == {("vector", 0), ("vector", 1), ("variadic", 0), ("variadic", 1)}


def test_reduce_partials_consistent():
    """Level-p partials of the oracle, reduced again with the level list shifted by p, give the
    same final value (tree separability)."""
    rng = np.random.default_rng(5)
    elem = QType(8, 8)
    lv = [QType(9 + l, 8 - l, True, l % 7, 0) for l in range(8)]
    for n in (512, 1000, 777):
        x = rng.integers(elem.raw_min, elem.raw_max + 1, size=n).astype(np.int32)
        plan = reduce_plan(elem, lv, n)
        full, parts = oracle.reduce_batched(plan, x, 1, partial_level=3)
        p2 = reduce_plan(lv[2], lv[3:], len(parts))
        again, _ = oracle.reduce_batched(p2, parts.astype(np.int32), 1)
        assert int(full[0]) == int(again[0])


def test_tfcmul_matches_reference():
    cases = load_cases("ref_tf.json")
    for c in cases:
        xr, xi, yr, yi = (qt(c[k]) for k in ("x_re", "x_im", "y_re", "y_im"))
        tags = tf_tags(c["tags"])
        r = resolve_tf(xr, xi, yr, yi, tags)
        plan = tfcmul_plan(xr, xi, yr, yi, tags, 1)
        wx = max(xr.word_bytes, xi.word_bytes); wy = max(yr.word_bytes, yi.word_bytes)
        for inp, out in zip(c["inputs"], c["outputs"]):
            assert out[2:7] == type_tuple(r["AB"]) and out[7:12] == type_tuple(r["BC"]), (c, out)
            x = np.array(inp[0:2], dtype=oracle.np_dtype(wx)); y = np.array(inp[2:4], dtype=oracle.np_dtype(wy))
            got = oracle.tfcmul_batched(plan, x, y, 1)
            assert [int(got[0]), int(got[1])] == out[0:2], (c["tags"], inp, out, got)
    # SURVEY 2.5 golden: x=(37,-22), y=(-51,19) -> re -44 (11.3), im 29 (12.2)
    assert cases[0]["inputs"][0] == [37, -22, -51, 19] and cases[0]["outputs"][0][0:2] == [-44, 29]


def test_basic_complex_mul_matches_reference():
    """Qmul<BasicComplexMul<...>>(x, y) and the default Qmul(x, y) on complex scalars (QuBLAS.h:3422-3446)."""
    import struct
    cases = load_cases("ref_bcm.json")
    for c in cases:
        xr, xi, yr, yi = (qt(c[k]) for k in ("x_re", "x_im", "y_re", "y_im"))
        r = resolve_bcm(xr, xi, yr, yi, tf_tags(c["tags"]))
        p = abi.qb_tfcmul_plan()
        p.x_re, p.x_im, p.y_re, p.y_im = (abi.desc(t) for t in (xr, xi, yr, yi))
        p.bcm = abi.bcm_types(r); p.product = abi.QB_CPROD_BASIC; p.len = 1
        wx = max(xr.word_bytes, xi.word_bytes); wy = max(yr.word_bytes, yi.word_bytes)
        for inp, out in zip(c["inputs"], c["outputs"]):
            assert out[2:7] == type_tuple(r["acbd"]) and out[7:12] == type_tuple(r["adbc"])
            got = oracle.tfcmul_batched(p, np.array(inp[:2], dtype=oracle.np_dtype(wx)), np.array(inp[2:], dtype=oracle.np_dtype(wy)), 1)
            assert [int(got[0]), int(got[1])] == out[:2], (c["tags"], inp)
    # BasicComplexMul of x=(37,-22), y=(-51,19) in 4.4 types -> re -91, im 113
    t44 = QType(4, 4)
    r = resolve_bcm(t44, t44, t44, t44, None)
    p = abi.qb_tfcmul_plan()
    p.x_re = p.x_im = p.y_re = p.y_im = abi.desc(t44); p.bcm = abi.bcm_types(r); p.product = abi.QB_CPROD_BASIC; p.len = 1
    got = oracle.tfcmul_batched(p, np.array([37, -22], dtype=np.int32), np.array([-51, 19], dtype=np.int32), 1)
    assert [int(got[0]), int(got[1])] == [-91, 113]


def _rc_plan_args(c):
    """(x_re, x_im, y_re, y_im, product, key of the imaginary product type) of a real-by-complex case."""
    r, cre, cim = qt(c["r"]), qt(c["c_re"]), qt(c["c_im"])
    return (r, None, cre, cim, "real_a", "ad") if c["real_first"] else (cre, cim, r, None, "real_b", "bc")


def test_real_complex_products_match_reference():
    """real (x) complex Qmul in both operand orders, every form of its target arguments (QuBLAS.h:3600-3637), and the
    mixed dot product = variadic complex Qreduce over such products: oracle == reference header."""
    from qublas_b200 import ops
    from qublas_b200.qtypes import resolve_rc
    from tests.helpers import rc_targs
    n_mul = n_dot = 0
    for c in load_cases("ref_rc.json"):
        if c["kind"] == "rc" and c["op"] == "mul":
            xr, xi, yr, yi, product, imk = _rc_plan_args(c)
            r = resolve_rc(c["real_first"], qt(c["r"]), qt(c["c_re"]), qt(c["c_im"]), rc_targs(c["targs"]))
            p = abi.qb_tfcmul_plan()
            p.x_re, p.x_im, p.y_re, p.y_im = (abi.desc(t) for t in (xr, xi or xr, yr, yi or yr))
            p.bcm = abi.bcm_types(r); p.product = abi.QB_CPROD_REAL_A if c["real_first"] else abi.QB_CPROD_REAL_B; p.len = 1
            for inp, out in zip(c["inputs"], c["outputs"]):
                assert out[2:7] == type_tuple(r["ac"]) and out[7:12] == type_tuple(r[imk]), c["targs"]
                real, cplx = np.array(inp[:1], dtype=np.int32), np.array(inp[1:], dtype=np.int32)
                got = oracle.tfcmul_batched(p, real if c["real_first"] else cplx, cplx if c["real_first"] else real, 1)
                assert [int(got[0]), int(got[1])] == out[:2], (c["targs"], inp)
                n_mul += 1
        elif c["kind"] == "rcgemul":
            K = c["K"]
            xr, xi, yr, yi, product, imk = _rc_plan_args(c)
            r = resolve_rc(c["real_first"], qt(c["r"]), qt(c["c_re"]), qt(c["c_im"]), rc_targs(c["targs"]))
            lv = [qt(t) for t in c["levels"]]
            o_re = tree_out_type(K, r["ac"], lv, TREE_VARIADIC); o_im = tree_out_type(K, r[imk], lv, TREE_VARIADIC)
            plan = ops.make_cgemul_plan(xr, xi, yr, yi, o_re, o_im, 1, 1, K, tf_tags=rc_targs(c["targs"]), add_re=lv, form=TREE_VARIADIC, product=product)
            for inp, out in zip(c["inputs"], c["outputs"]):
                v = np.array(inp, dtype=np.int32).reshape(K, 3)
                real, cplx = np.ascontiguousarray(v[:, 0]), np.ascontiguousarray(v[:, 1:3]).reshape(-1)
                assert out[2:7] == type_tuple(o_re) and out[7:12] == type_tuple(o_im)
                got = oracle.cgemul_batched(plan, real if c["real_first"] else cplx, cplx if c["real_first"] else real, 1)
                assert [int(got[0]), int(got[1])] == out[0:2], (K, c["real_first"], out, got)
                n_dot += 1
    assert n_mul >= 400 and n_dot >= 48


def test_gemul_matches_reference_composition():
    cases = load_cases("ref_gemul.json")
    for c in cases:
        a, b, mul, cc = (qt(c[k]) for k in ("a", "b", "mul", "c"))
        lv = [qt(t) for t in c["levels"]]
        M, N, K = c["M"], c["N"], c["K"]
        plan = gemul_plan(a, b, mul, cc, lv, M, N, K, trans_a=c["trans_a"], form=form_id(c["form"]))
        for inp, out in zip(c["inputs"], c["outputs"]):
            A = np_words(inp[:M * K], a); B = np_words(inp[M * K:], b)
            got = oracle.gemul_batched(plan, A, B, 1)
            assert [int(v) for v in got] == out, (c["M"], c["N"], c["K"], c["form"])
    # readme.md:80-87 on m1 = 1..16: SURVEY 3.5 goldens
    assert cases[0]["outputs"][0] == [192, 448, 192, 448, 448, 448, 0, 0, 192, 0, 0, 0, 448, 0, 0, 0]
    assert cases[1]["outputs"][0] == [64, 64, 128, 448, 448, 320, 448, 0, 448, 0, 192, 256, 64, 192, 256, 384]


def test_cgemul_matches_reference_composition():
    cases = load_cases("ref_cgemul.json")
    for c in cases:
        xr, xi, yr, yi = (qt(c[k]) for k in ("x_re", "x_im", "y_re", "y_im"))
        tags = tf_tags(c["tags"]); lv = [qt(t) for t in c["levels"]]
        K = c["K"]
        r = resolve_tf(xr, xi, yr, yi, tags)
        o_re = tree_out_type(K, r["AB"], lv, TREE_VARIADIC); o_im = tree_out_type(K, r["BC"], lv, TREE_VARIADIC)
        # 1 x K times K x 1 complex Qgemul; C type = the reduction's own result type (identity conversion)
        plan = cgemul_plan(xr, xi, yr, yi, tags, lv, lv, o_re, o_im, 1, 1, K, trans_a=False, form=TREE_VARIADIC)
        for inp, out in zip(c["inputs"], c["outputs"]):
            v = np.array(inp, dtype=np.int32).reshape(K, 4)
            A = np.ascontiguousarray(v[:, 0:2]).reshape(-1); B = np.ascontiguousarray(v[:, 2:4]).reshape(-1)
            assert out[2:7] == type_tuple(o_re) and out[7:12] == type_tuple(o_im)
            got = oracle.cgemul_batched(plan, A, B, 1)
            assert [int(got[0]), int(got[1])] == out[0:2], (K, out, got)


def test_readme_reductions():
    """readme.md:57,61,64 (SURVEY 8(c) goldens): red1 = 0, red2 raw 7 @ 6.-3 (=56), red3 = 34."""
    type1 = QType(6, 3, True, 5, 1); type2 = QType(6, -3)
    v1 = QType(4, 8)                                  # Qu<dim<4>, intBits<4>> -> fracBits default 8
    x = np.array([int(v * 256) for v in (1.0, 2.0, 3.0, 4.0)], dtype=np.int32)
    # values 1..4 in a 4.8 type: 4.0 saturates at construction, but raw inputs here are what the readme stores
    x = np.minimum(x, v1.raw_max)
    r1, _ = oracle.reduce_batched(reduce_plan(v1, [type2], 4), x, 1)
    assert int(r1[0]) * 2 ** 3 == 0
    m1 = np.array([8 * v for v in range(1, 17)], dtype=np.int32)
    r2, _ = oracle.reduce_batched(reduce_plan(type1, [type1, type2], 16), m1, 1)
    assert int(r2[0]) == 7
    # Qreduce<type1>(q1, q2, q1, q2), q1 = 1.0 (raw 8 @ 6.3), q2 = 16 (raw 2 @ 6.-3): mixed-type variadic
    a = oracle.qadd(8, type1, 2, type2, type1)
    r3 = oracle.qadd(a, type1, a, type1, type1)
    assert r3 / 8 == 34


def test_from_double_matches_reference():
    """Qu_s(double) (QuBLAS.h:2387-2393): the round-1 goldens, and the round-2 goldens for 64-bit-wide targets with
    |raw| around 2^62..2^64 (outside the reference's defect band, see helpers.fd64_defect_band)."""
    import struct
    from tests.helpers import fd64_defect_band
    n = skipped = 0
    for name in ("ref_fromdouble.json", "ref_fromdouble64.json"):
        for c in load_cases(name):
            t = qt(c["to"])
            for i, o in zip(c["inputs"], c["outputs"]):
                v = struct.unpack("<d", struct.pack("<q", i[0]))[0]
                if c.get("big") and fd64_defect_band(v, t):
                    skipped += 1
                    continue
                assert oracle.from_double(v, t) == o[0], (c["to"], v)
                n += 1
    assert n >= 4500 and skipped < 1300


def test_anus_poly_and_approx_match_reference():
    """ANUS::Qpoly / Qapprox (reference QuBLAS.h:4829-4897): the oracle's restatement against the reference's own outputs
    (tests/golden/ref_anus.txt, generated by oracle/ref_harness_gen.py --text from the unmodified header)."""
    from qublas_b200 import ops
    from tests.helpers import anus_golden
    x_t, exprs, xs, want = anus_golden()
    for name, segs in exprs.items():
        plan = ops.make_poly_plan(x_t, segs, len(xs), approx=name in ("q", "r"))
        got = oracle.poly_batched(plan, xs)
        assert np.array_equal(got.astype(np.int64), want[name]), name
