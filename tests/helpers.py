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
