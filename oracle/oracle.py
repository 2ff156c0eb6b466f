"""ctypes loader for the CPU oracle (oracle/qb_oracle.c).  TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  The product package (qublas_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

_LIB = None


# The checker does not import the product: plans arrive as ctypes structures laid out like include/qb200.h (whoever built
# them) and are passed by address; the few PODs the oracle has to build itself are mirrored here.
class qb_qdesc(C.Structure):
    _fields_ = [("int_bits", C.c_int16), ("frac_bits", C.c_int16), ("is_signed", C.c_uint8),
                ("qu_mode", C.c_uint8), ("of_mode", C.c_uint8), ("word_bytes", C.c_uint8)]


def _desc(t, word_bytes=0):
    """qb_qdesc of a type given as an object with .i .f .s .q .o (e.g. qtypes.QType) or as a qb_qdesc-like structure."""
    if hasattr(t, "int_bits"):
        return qb_qdesc(t.int_bits, t.frac_bits, t.is_signed, t.qu_mode, t.of_mode, t.word_bytes)
    return qb_qdesc(t.i, t.f, 1 if t.s else 0, t.q, t.o, word_bytes)


def _word_bytes(d):
    """storage word of a qb_qdesc: the explicit one, else the reference's natural word (QuBLAS.h:353)."""
    return d.word_bytes or (4 if 1 + d.int_bits + d.frac_bits <= 32 else 8)


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
        i64, u64, u32, vp, qd = C.c_int64, C.c_uint64, C.c_uint32, C.c_void_p, C.POINTER(qb_qdesc)
        L.qbo_conv_i64.restype = i64; L.qbo_conv_i64.argtypes = [i64, C.c_int, qd]
        L.qbo_conv_i128.restype = i64; L.qbo_conv_i128.argtypes = [i64, u64, C.c_int, qd]
        for f in (L.qbo_qmul_i64, L.qbo_qadd_i64, L.qbo_qsub_i64):
            f.restype = i64; f.argtypes = [i64, qd, i64, qd, qd]
        L.qbo_reduce_out_type.argtypes = [vp, qd]
        L.qbo_reduce_batched.argtypes = [vp, vp, vp, vp, u32, u64]
        L.qbo_gemul_batched.argtypes = [vp, vp, vp, vp, u64]
        L.qbo_cgemul_batched.argtypes = [vp, vp, vp, vp, u64]
        L.qbo_tfcmul_batched.argtypes = [vp, vp, vp, vp, u64]
        L.qbo_convert_batched.argtypes = [vp, vp, vp, u64]
        L.qbo_elementwise_batched.argtypes = [vp, vp, vp, vp, u64]
        L.qbo_poly_batched.argtypes = [vp, vp, vp, u64]
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


def conv(raw, from_frac, to):
    return load().qbo_conv_i64(int(raw), int(from_frac), C.byref(_desc(to)))


def conv128(raw, from_frac, to):
    raw = int(raw)
    lo = raw & ((1 << 64) - 1)
    hi = raw >> 64
    return load().qbo_conv_i128(hi, lo, int(from_frac), C.byref(_desc(to)))


def qmul(a, ta, b, tb, to):
    return load().qbo_qmul_i64(int(a), C.byref(_desc(ta)), int(b), C.byref(_desc(tb)), C.byref(_desc(to)))


def qadd(a, ta, b, tb, to):
    return load().qbo_qadd_i64(int(a), C.byref(_desc(ta)), int(b), C.byref(_desc(tb)), C.byref(_desc(to)))


def qsub(a, ta, b, tb, to):
    return load().qbo_qsub_i64(int(a), C.byref(_desc(ta)), int(b), C.byref(_desc(tb)), C.byref(_desc(to)))


def reduce_batched(plan, x: np.ndarray, batch: int, partial_level=None):
    """Returns (out, partials|None); out dtype = natural word of the result type."""
    ot = qb_qdesc()
    assert load().qbo_reduce_out_type(C.byref(plan), C.byref(ot)) == 0
    out = np.zeros(batch, dtype=np_dtype(_word_bytes(ot)))
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


def gemul_batched(plan, A, B, batch):
    wb = _word_bytes(plan.c)
    Cc = np.zeros(batch * plan.M * plan.N if plan.stride_c == plan.M * plan.N else (batch - 1) * plan.stride_c + plan.M * plan.N,
                  dtype=np_dtype(wb))
    A = np.ascontiguousarray(A); B = np.ascontiguousarray(B)
    rc = load().qbo_gemul_batched(C.byref(plan), _ptr(A), _ptr(B), _ptr(Cc), batch)
    assert rc == 0, rc
    return Cc


def cword_bytes(re, im):
    return max(_word_bytes(re), _word_bytes(im))


def cgemul_batched(plan, A, B, batch):
    wb = cword_bytes(plan.c_re, plan.c_im)
    Cc = np.zeros(2 * batch * plan.M * plan.N, dtype=np_dtype(wb))
    A = np.ascontiguousarray(A); B = np.ascontiguousarray(B)
    rc = load().qbo_cgemul_batched(C.byref(plan), _ptr(A), _ptr(B), _ptr(Cc), batch)
    assert rc == 0, rc
    return Cc


def tfcmul_batched(plan, x, y, batch):
    # enum qb_cproduct (include/qb200.h): 0 = TF, 1 = BASIC, 2 = REAL_A, 3 = REAL_B
    re_t, im_t = {0: (plan.tf.AB, plan.tf.BC), 1: (plan.bcm.acbd, plan.bcm.adbc),
                  2: (plan.bcm.ac, plan.bcm.ad), 3: (plan.bcm.ac, plan.bcm.bc)}[plan.product]
    wb = cword_bytes(re_t, im_t)
    out = np.zeros(2 * batch * plan.len, dtype=np_dtype(wb))
    x = np.ascontiguousarray(x); y = np.ascontiguousarray(y)
    rc = load().qbo_tfcmul_batched(C.byref(plan), _ptr(x), _ptr(y), _ptr(out), batch)
    assert rc == 0, rc
    return out


def convert_batched(plan, x, batch):
    out = np.zeros(batch * plan.len, dtype=np_dtype(_word_bytes(plan.to)))
    x = np.ascontiguousarray(x)
    rc = load().qbo_convert_batched(C.byref(plan), _ptr(x), _ptr(out), batch)
    assert rc == 0, rc
    return out


def synth_fill(count, word_bytes, lo, hi, seed, stream_id, first_index=0):
    out = np.zeros(count, dtype=np_dtype(word_bytes))
    rc = load().qbo_synth_fill(_ptr(out), count, word_bytes, lo, hi, seed, stream_id, first_index)
    assert rc == 0
    return out


def from_double(v, t):
    return load().qbo_from_double(float(v), C.byref(_desc(t)))


def quantize_f64(t, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros(x.size, dtype=np_dtype(_word_bytes(_desc(t))))
    assert load().qbo_quantize_f64(C.byref(_desc(t)), _ptr(x), _ptr(out), x.size) == 0
    return out


def dequantize_f64(t, w):
    w = np.ascontiguousarray(w)
    out = np.zeros(w.size, dtype=np.float64)
    assert load().qbo_dequantize_f64(C.byref(_desc(t)), _ptr(w), _ptr(out), w.size) == 0
    return out


def error_stats(t, w, ref):
    w = np.ascontiguousarray(w); ref = np.ascontiguousarray(ref, dtype=np.float64)
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    assert load().qbo_error_stats(C.byref(_desc(t)), _ptr(w), _ptr(ref), w.size, C.byref(a), C.byref(b), C.byref(c)) == 0
    return a.value, b.value, c.value


def elementwise(plan, a, b, batch=1):
    out = np.zeros(batch * plan.len, dtype=np_dtype(_word_bytes(plan.to)))
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    assert load().qbo_elementwise_batched(C.byref(plan), _ptr(a), _ptr(b), _ptr(out), batch) == 0
    return out


def poly_batched(plan, x, batch=1):
    """ANUS::Qpoly / Qapprox over a tensor; plan = a qb_poly_plan-shaped ctypes structure (include/qb200.h)."""
    ot = plan.x if plan.approx else plan.seg[0].coef[0].type
    out = np.zeros(batch * plan.len, dtype=np_dtype(_word_bytes(ot)))
    x = np.ascontiguousarray(x)
    assert load().qbo_poly_batched(C.byref(plan), _ptr(x), _ptr(out), batch) == 0
    return out
