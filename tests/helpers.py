"""Shared test helpers: fixture loading and plan construction from golden-case dicts."""
import json
import os

import numpy as np

from qublas_b200 import abi
from qublas_b200.qtypes import QType, TREE_VARIADIC, TREE_VECTOR, resolve_tf

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_cases(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)["cases"]


def qt(d):
    return QType.from_dict(d)


def form_id(s):
    return TREE_VARIADIC if s == "variadic" else TREE_VECTOR


def np_words(vals, t: QType):
    return np.array(vals, dtype=np.int32 if t.word_bytes == 4 else np.int64)


def reduce_plan(elem: QType, levels, n, form=TREE_VECTOR):
    p = abi.qb_reduce_plan()
    p.elem = abi.desc(elem)
    p.levels = abi.levels(levels)
    p.len = n
    p.tree_form = form
    return p


def gemul_plan(a, b, mul, c, levels, M, N, K, trans_a=False, trans_b=False, form=TREE_VECTOR):
    p = abi.qb_gemul_plan()
    p.a, p.b, p.mul, p.c = abi.desc(a), abi.desc(b), abi.desc(mul), abi.desc(c)
    p.add = abi.levels(levels)
    p.M, p.N, p.K = M, N, K
    p.trans_a, p.trans_b, p.tree_form = int(trans_a), int(trans_b), form
    p.stride_a, p.stride_b, p.stride_c = M * K, K * N, M * N
    return p


def cgemul_plan(a_re, a_im, b_re, b_im, tags, lv_re, lv_im, c_re, c_im, M, N, K, trans_a=False, trans_b=False,
                form=TREE_VARIADIC):
    p = abi.qb_cgemul_plan()
    p.a_re, p.a_im, p.b_re, p.b_im = (abi.desc(t) for t in (a_re, a_im, b_re, b_im))
    p.tf = abi.tfc_types(resolve_tf(a_re, a_im, b_re, b_im, tags))
    p.add_re, p.add_im = abi.levels(lv_re), abi.levels(lv_im)
    p.c_re, p.c_im = abi.desc(c_re), abi.desc(c_im)
    p.M, p.N, p.K = M, N, K
    p.trans_a, p.trans_b, p.tree_form = int(trans_a), int(trans_b), form
    p.stride_a, p.stride_b, p.stride_c = M * K, K * N, M * N
    return p


def tfcmul_plan(x_re, x_im, y_re, y_im, tags, n):
    p = abi.qb_tfcmul_plan()
    p.x_re, p.x_im, p.y_re, p.y_im = (abi.desc(t) for t in (x_re, x_im, y_re, y_im))
    p.tf = abi.tfc_types(resolve_tf(x_re, x_im, y_re, y_im, tags))
    p.len = n
    return p


def tf_tags(d):
    return {k: qt(v) for k, v in d.items()}


def rc_targs(t):
    """Merger arguments of a ref_rc.json case in the form qtypes.resolve_rc takes."""
    if t["mode"] == "none":
        return None
    if t["mode"] == "one":
        return qt(t["t"])
    if t["mode"] == "pair":
        return (qt(t["re"]), qt(t["im"]))
    return {"realT": qt(t["re"]), "imagT": qt(t["im"])}


def fd64_defect_band(v, t: QType):
    """True when a double -> Qu conversion falls into the reference's defect band: with a SAT OfMode, a rounded raw value
    of magnitude in [2^63, 2^64) is compared against the bounds through the low 64-bit word only (SURVEY.md 2.3 item 4,
    QuBLAS.h:1792) and comes back as its truncated two's-complement word instead of the saturated bound.  The library and
    the oracle implement the documented behaviour (saturate); the band is excluded from the parity domain."""
    if t.o == 3:                                  # WRP::TCPL keeps the low bits either way: reference == ideal model
        return False
    raw = abs(v) * 2.0 ** t.f
    return 2.0 ** 63 * (1 - 2.0 ** -40) <= raw <= 2.0 ** 64 * (1 + 2.0 ** -40)


def anus_golden():
    """tests/golden/ref_anus.txt = stdout of tests/cpp/anus_example.cpp compiled against the REFERENCE header: per x raw value
    the raws of six ANUS::Qpoly / Qapprox expressions.  Returns (x type, {name: segments}, x raws, {name: expected raws})."""
    from oracle import oracle
    from qublas_b200.qtypes import RND_POS_INF, SAT_ZERO, TRN_SMGN
    x_t = QType(4, 8)
    c_t = QType(3, 6, True, RND_POS_INF)
    w_t = QType(2, 10, True, TRN_SMGN, SAT_ZERO)
    a = [(c_t, oracle.from_double(v, c_t)) for v in (0.5, -1.25, 0.375, 2.0)]
    b = [(w_t, oracle.from_double(v, w_t)) for v in (-0.3333, 0.75, -0.0625)]
    exprs = {
        "p1": [(0.0, [a[0]])],
        "p2": [(0.0, [a[0], a[1]])],
        "p3": [(0.0, [a[0], a[1], a[2]])],
        "p4": [(0.0, [b[0], b[1], b[2], a[3]])],
        "q": [(-2.0, [a[0], a[1]]), (0.5, [b[0], b[1], b[2]]), (3.0, [a[3], a[2]])],
        "r": [(1.0, [b[1]]), (100.0, [a[1], b[2]])],
    }
    xs, want = [], {k: [] for k in exprs}
    with open(os.path.join(GOLD, "ref_anus.txt")) as f:
        for line in f:
            t = line.split()
            xs.append(int(t[0]))
            for k, v in zip(("p1", "p2", "p3", "p4", "q", "r"), t[1:]):
                want[k].append(int(v.split(":")[0]))
    return x_t, exprs, np.array(xs, dtype=np.int32), {k: np.array(v, dtype=np.int64) for k, v in want.items()}
