"""ctypes loader for the CPU oracle (oracle/qb_oracle.c).  TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  The product package (qublas_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from qublas_b200 import abi  # noqa: E402  (POD struct definitions only)
from qublas_b200.qtypes import QType, tree_out_type  # noqa: E402

_LIB = None


def build():
    so = os.path.join(HERE, "_build", "libqb_oracle.so")
    src = os.path.join(HERE, "qb_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, src, "-lm"])
    return so


def load():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        i64, u64, u32, vp, qd = C.c_int64, C.c_uint64, C.c_uint32, C.c_void_p, C.POINTER(abi.qb_qdesc)
        L.qbo_conv_i64.restype = i64; L.qbo_conv_i64.argtypes = [i64, C.c_int, qd]
        L.qbo_conv_i128.restype = i64; L.qbo_conv_i128.argtypes = [i64, u64, C.c_int, qd]
        for f in (L.qbo_qmul_i64, L.qbo_qadd_i64, L.qbo_qsub_i64):
            f.restype = i64; f.argtypes = [i64, qd, i64, qd, qd]
        L.qbo_reduce_out_type.argtypes = [C.POINTER(abi.qb_reduce_plan), qd]
        L.qbo_reduce_batched.argtypes = [C.POINTER(abi.qb_reduce_plan), vp, vp, vp, u32, u64]
        L.qbo_gemul_batched.argtypes = [C.POINTER(abi.qb_gemul_plan), vp, vp, vp, u64]
        L.qbo_cgemul_batched.argtypes = [C.POINTER(abi.qb_cgemul_plan), vp, vp, vp, u64]
        L.qbo_tfcmul_batched.argtypes = [C.POINTER(abi.qb_tfcmul_plan), vp, vp, vp, u64]
        L.qbo_convert_batched.argtypes = [C.POINTER(abi.qb_convert_plan), vp, vp, u64]
        L.qbo_elementwise_batched.argtypes = [C.POINTER(abi.qb_ew_plan), vp, vp, vp, u64]
        L.qbo_synth_fill.argtypes = [vp, u64, u32, i64, i64, u64, u64, u64]
        dp = C.POINTER(C.c_double)
        L.qbo_from_double.restype = i64; L.qbo_from_double.argtypes = [C.c_double, qd]
        L.qbo_quantize_f64.argtypes = [qd, vp, vp, u64]
        L.qbo_dequantize_f64.argtypes = [qd, vp, vp, u64]
        L.qbo_error_stats.argtypes = [qd, vp, vp, u64, dp, dp, dp]
        L.qbo_synth_word.restype = i64; L.qbo_synth_word.argtypes = [u64, u64, u64, i64, i64]
        _LIB = L
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def np_dtype(word_bytes):
    return {1: np.int8, 2: np.int16, 4: np.int32, 8: np.int64}[word_bytes]


def conv(raw, from_frac, to: QType):
    return load().qbo_conv_i64(int(raw), int(from_frac), C.byref(abi.desc(to)))


def conv128(raw, from_frac, to: QType):
    raw = int(raw)
    lo = raw & ((1 << 64) - 1)
    hi = raw >> 64
    return load().qbo_conv_i128(hi, lo, int(from_frac), C.byref(abi.desc(to)))


def qmul(a, ta, b, tb, to):
    return load().qbo_qmul_i64(int(a), C.byref(abi.desc(ta)), int(b), C.byref(abi.desc(tb)), C.byref(abi.desc(to)))


def qadd(a, ta, b, tb, to):
    return load().qbo_qadd_i64(int(a), C.byref(abi.desc(ta)), int(b), C.byref(abi.desc(tb)), C.byref(abi.desc(to)))


def qsub(a, ta, b, tb, to):
    return load().qbo_qsub_i64(int(a), C.byref(abi.desc(ta)), int(b), C.byref(abi.desc(tb)), C.byref(abi.desc(to)))


def reduce_batched(plan: abi.qb_reduce_plan, x: np.ndarray, batch: int, partial_level=None):
    """Returns (out, partials|None); out dtype = natural word of the result type."""
    elem = abi.qtype_of(plan.elem)
    lv = [abi.qtype_of(plan.levels.lv[i]) for i in range(plan.levels.n)]
    ot = tree_out_type(int(plan.len), elem, lv, int(plan.tree_form))
    out = np.zeros(batch, dtype=np_dtype(ot.word_bytes))
    partials = None
    pl = 0
    if partial_level is not None:
        per = int(plan.len)
        for _ in range(partial_level):
            per = (per + 1) // 2
        partials = np.zeros(batch * per, dtype=np.int64)
        pl = partial_level
    x = np.ascontiguousarray(x)
    rc = load().qbo_reduce_batched(C.byref(plan), _ptr(x), _ptr(out), _ptr(partials) if partials is not None else None, pl, batch)
    assert rc == 0, rc
    return out, partials


def gemul_batched(plan: abi.qb_gemul_plan, A, B, batch):
    c = abi.qtype_of(plan.c)
    wb = plan.c.word_bytes or c.word_bytes
    Cc = np.zeros(batch * plan.M * plan.N if plan.stride_c == plan.M * plan.N else (batch - 1) * plan.stride_c + plan.M * plan.N,
                  dtype=np_dtype(wb))
    A = np.ascontiguousarray(A); B = np.ascontiguousarray(B)
    rc = load().qbo_gemul_batched(C.byref(plan), _ptr(A), _ptr(B), _ptr(Cc), batch)
    assert rc == 0, rc
    return Cc


def cword_bytes(re: abi.qb_qdesc, im: abi.qb_qdesc):
    f = lambda d: d.word_bytes or abi.qtype_of(d).word_bytes
    return max(f(re), f(im))


def cgemul_batched(plan: abi.qb_cgemul_plan, A, B, batch):
    wb = cword_bytes(plan.c_re, plan.c_im)
    Cc = np.zeros(2 * batch * plan.M * plan.N, dtype=np_dtype(wb))
    A = np.ascontiguousarray(A); B = np.ascontiguousarray(B)
    rc = load().qbo_cgemul_batched(C.byref(plan), _ptr(A), _ptr(B), _ptr(Cc), batch)
    assert rc == 0, rc
    return Cc


def tfcmul_batched(plan: abi.qb_tfcmul_plan, x, y, batch):
    re_t, im_t = {abi.QB_CPROD_TF: (plan.tf.AB, plan.tf.BC), abi.QB_CPROD_BASIC: (plan.bcm.acbd, plan.bcm.adbc),
                  abi.QB_CPROD_REAL_A: (plan.bcm.ac, plan.bcm.ad), abi.QB_CPROD_REAL_B: (plan.bcm.ac, plan.bcm.bc)}[plan.product]
    wb = cword_bytes(re_t, im_t)
    out = np.zeros(2 * batch * plan.len, dtype=np_dtype(wb))
    x = np.ascontiguousarray(x); y = np.ascontiguousarray(y)
    rc = load().qbo_tfcmul_batched(C.byref(plan), _ptr(x), _ptr(y), _ptr(out), batch)
    assert rc == 0, rc
    return out


def convert_batched(plan: abi.qb_convert_plan, x, batch):
    to = abi.qtype_of(plan.to)
    out = np.zeros(batch * plan.len, dtype=np_dtype(plan.to.word_bytes or to.word_bytes))
    x = np.ascontiguousarray(x)
    rc = load().qbo_convert_batched(C.byref(plan), _ptr(x), _ptr(out), batch)
    assert rc == 0, rc
    return out


def synth_fill(count, word_bytes, lo, hi, seed, stream_id, first_index=0):
    out = np.zeros(count, dtype=np_dtype(word_bytes))
    rc = load().qbo_synth_fill(_ptr(out), count, word_bytes, lo, hi, seed, stream_id, first_index)
    assert rc == 0
    return out


def from_double(v, t: QType):
    return load().qbo_from_double(float(v), C.byref(abi.desc(t)))


def quantize_f64(t: QType, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros(x.size, dtype=np_dtype(t.word_bytes))
    assert load().qbo_quantize_f64(C.byref(abi.desc(t)), _ptr(x), _ptr(out), x.size) == 0
    return out


def dequantize_f64(t: QType, w):
    w = np.ascontiguousarray(w)
    out = np.zeros(w.size, dtype=np.float64)
    assert load().qbo_dequantize_f64(C.byref(abi.desc(t)), _ptr(w), _ptr(out), w.size) == 0
    return out


def error_stats(t: QType, w, ref):
    w = np.ascontiguousarray(w); ref = np.ascontiguousarray(ref, dtype=np.float64)
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    assert load().qbo_error_stats(C.byref(abi.desc(t)), _ptr(w), _ptr(ref), w.size, C.byref(a), C.byref(b), C.byref(c)) == 0
    return a.value, b.value, c.value


def elementwise(plan: abi.qb_ew_plan, a, b, batch=1):
    to = abi.qtype_of(plan.to)
    out = np.zeros(batch * plan.len, dtype=np_dtype(plan.to.word_bytes or to.word_bytes))
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    assert load().qbo_elementwise_batched(C.byref(plan), _ptr(a), _ptr(b), _ptr(out), batch) == 0
    return out
