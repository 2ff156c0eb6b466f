"""Host-side mirror of the reference's operator interface for the hot path, over the C ABI.

Same names and argument meaning as the reference's template API:

    Qreduce<Args...>(tensor)                                   QuBLAS.h:5014-5018
    Qgemul<QgemulAddArgs<..>, QgemulMulArgs<..>, QgemulTransposedA<..>>(C, A, B)   readme.md:80-87
    Qmul<TFComplexMul<...>>(x, y) on complex tensors           QuBLAS.h:3510-3535, 4080-4085

Types are `QType`s instead of template arguments; tensors are raw two's-complement words in the
reference's layout (column-major, int32 for width <= 32 else int64, complex = AoS pairs), held
either in numpy arrays (host memory; the library stages them through the GPU) or in torch CUDA
tensors (device memory; runs on torch's current stream).  Errors surface as QbError, like the
reference's std::runtime_error.  Everything executes in libqb200.so's CUDA kernels — there is no
CPU path here.
"""
import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import abi
from .qtypes import (QType, TREE_VARIADIC, TREE_VECTOR, add_merge, mul_merge, resolve_bcm, resolve_rc, resolve_tf, tree_out_type)

try:  # torch is plumbing (device memory, streams); host-only use works without it
    import torch
except Exception:  # pragma: no cover
    torch = None

_NP = {1: np.int8, 2: np.int16, 4: np.int32, 8: np.int64}


def _is_torch(x):
    return torch is not None and isinstance(x, torch.Tensor)


def _tdtype(wb):
    return {1: torch.int8, 2: torch.int16, 4: torch.int32, 8: torch.int64}[wb]


def _exec_for(*tensors, flags=0) -> abi.qb_exec:
    dev = [t for t in tensors if _is_torch(t)]
    if dev:
        if len(dev) != len([t for t in tensors if t is not None]):
            raise ValueError("mix of host and device buffers")
        d = dev[0].device
        if d.type != "cuda" or any(t.device != d for t in dev):
            raise ValueError("all tensors must live on one CUDA device")
        return abi.qb_exec(abi.QB_MEM_DEVICE, d.index if d.index is not None else torch.cuda.current_device(),
                           torch.cuda.current_stream(d).cuda_stream, flags, 0)
    return abi.qb_exec(abi.QB_MEM_HOST, -1, None, flags, 0)


def _ptr(x):
    if _is_torch(x):
        if not x.is_contiguous():
            raise ValueError("tensors must be contiguous")
        return C.c_void_p(x.data_ptr())
    if not x.flags["C_CONTIGUOUS"]:
        raise ValueError("arrays must be contiguous")
    return x.ctypes.data_as(C.c_void_p)


def _check_words(x, wb, what):
    isz = x.element_size() if _is_torch(x) else x.dtype.itemsize
    kind_ok = (not x.dtype.is_floating_point) if _is_torch(x) else (x.dtype.kind == "i")
    if isz != wb or not kind_ok:
        raise ValueError("%s: expected %d-byte signed integer raw words" % (what, wb))


def _new_like(ref, n, wb):
    if _is_torch(ref):
        return torch.empty(n, dtype=_tdtype(wb), device=ref.device)
    return np.empty(n, dtype=_NP[wb])


def _wb(t: QType, word_bytes: int = 0):
    return word_bytes or t.word_bytes


def plan_info_reduce(plan):
    info = abi.qb_plan_info()
    abi.check(abi.load().qb200_reduce_plan_info(C.byref(plan), C.byref(info)))
    return info


def plan_info_gemul(plan):
    info = abi.qb_plan_info()
    abi.check(abi.load().qb200_gemul_plan_info(C.byref(plan), C.byref(info)))
    return info


def plan_info_cgemul(plan):
    info = abi.qb_plan_info()
    abi.check(abi.load().qb200_cgemul_plan_info(C.byref(plan), C.byref(info)))
    return info


# --------------------------------------------------------------------------------------------
def make_reduce_plan(elem: QType, levels: Optional[Sequence[QType]], length: int, form: int = TREE_VECTOR,
                     word_bytes: int = 0) -> abi.qb_reduce_plan:
    p = abi.qb_reduce_plan()
    p.elem = abi.desc(elem, word_bytes)
    p.levels = abi.levels(levels)
    p.len = length
    p.tree_form = form
    return p


def Qreduce(x, elem: QType, levels: Optional[Sequence[QType]] = None, *, length: Optional[int] = None, batch: int = 1,
            form: int = TREE_VECTOR, partial_level: Optional[int] = None, word_bytes: int = 0, flags: int = 0,
            plan: Optional[abi.qb_reduce_plan] = None, out=None):
    """Tree reduction of `batch` vectors of `length` elements (default: the whole buffer is one vector).

    Returns (out, out_type) or (out, out_type, partials) when partial_level is given.  `plan` (from
    make_reduce_plan) and `out` may be passed in to keep repeated calls allocation-free.
    """
    total = x.numel() if _is_torch(x) else x.size
    if length is None:
        length = total // batch
    if length * batch != total:
        raise ValueError("buffer size does not match batch * length")
    _check_words(x, _wb(elem, word_bytes), "x")
    if plan is None:
        plan = make_reduce_plan(elem, levels, length, form, word_bytes)
    ot = tree_out_type(length, elem, list(levels or []), form)
    if out is None:
        out = _new_like(x, batch, ot.word_bytes)
    parts = None
    if partial_level is not None:
        per = length
        for _ in range(partial_level):
            per = (per + 1) // 2
        parts = _new_like(x, batch * per, 8)
    ex = _exec_for(x, flags=flags)
    abi.check(abi.load().qb200_reduce_batched(C.byref(plan), _ptr(x), _ptr(out), _ptr(parts) if parts is not None else None,
                                              partial_level or 0, batch, C.byref(ex)))
    return (out, ot) if parts is None else (out, ot, parts)


# --------------------------------------------------------------------------------------------
def make_gemul_plan(a: QType, b: QType, c: QType, M: int, N: int, K: int, *, mul=None, add: Optional[Sequence[QType]] = None,
                    trans_a: bool = False, trans_b: bool = False, form: int = TREE_VECTOR,
                    word_bytes=(0, 0, 0)) -> abi.qb_gemul_plan:
    """mul: QgemulMulArgs — a full QType, a dict of partial tags (i/f/s/q/o/full) or None (MulMerger defaults)."""
    p = abi.qb_gemul_plan()
    mt = mul if isinstance(mul, QType) else mul_merge(a, b, mul)
    p.a, p.b, p.c = abi.desc(a, word_bytes[0]), abi.desc(b, word_bytes[1]), abi.desc(c, word_bytes[2])
    p.mul = abi.desc(mt)
    p.add = abi.levels(add)
    p.M, p.N, p.K = M, N, K
    p.trans_a, p.trans_b, p.tree_form = int(trans_a), int(trans_b), form
    p.stride_a, p.stride_b, p.stride_c = M * K, K * N, M * N
    return p


def _numel(x):
    return x.numel() if _is_torch(x) else x.size


def _check_operand(x, desc, words_per_elem, batch, stride, per_instance, what):
    """Raw-word buffer of a batched operand: word size of the plan's type, long enough for (batch-1)*stride + one instance."""
    wb = desc.word_bytes or abi.qtype_of(desc).word_bytes
    _check_words(x, wb, what)
    need = words_per_elem * ((batch - 1) * stride + per_instance) if batch > 0 else 0
    if _numel(x) < need:
        raise ValueError("%s: buffer holds %d words, the plan needs %d" % (what, _numel(x), need))


def Qgemul(Cbuf, A, B, plan: abi.qb_gemul_plan, batch: int = 1, flags: int = 0):
    """Qgemul<AddArgs, MulArgs, TransposedA>(C, A, B): out-parameter first, like the reference."""
    lib = abi.load()
    _check_operand(A, plan.a, 1, batch, plan.stride_a, plan.M * plan.K, "A")
    _check_operand(B, plan.b, 1, batch, plan.stride_b, plan.K * plan.N, "B")
    _check_operand(Cbuf, plan.c, 1, batch, plan.stride_c, plan.M * plan.N, "C")
    ex = _exec_for(Cbuf, A, B, flags=flags)
    abi.check(lib.qb200_gemul_batched(C.byref(plan), _ptr(A), _ptr(B), _ptr(Cbuf), batch, C.byref(ex)))
    return Cbuf


def make_cgemul_plan(a_re: QType, a_im: QType, b_re: QType, b_im: QType, c_re: QType, c_im: QType, M: int, N: int, K: int, *,
                     tf_tags=None, add_re: Optional[Sequence[QType]] = None, add_im: Optional[Sequence[QType]] = None,
                     trans_a=False, trans_b=False, form: int = TREE_VARIADIC, product: str = "tf") -> abi.qb_cgemul_plan:
    """product: "tf" = TFComplexMul<tf_tags> (3 mult / 5 add), "basic" = BasicComplexMul<tf_tags> (4 mult / 2 add),
    "real_a" / "real_b" = real (x) complex Qmul<tf_tags> with A resp. B a REAL tensor (its *_im type is ignored;
    tf_tags as in qtypes.resolve_rc)."""
    p = abi.qb_cgemul_plan()
    p.a_re, p.a_im, p.b_re, p.b_im = (abi.desc(t) for t in (a_re, a_im if a_im is not None else a_re, b_re, b_im if b_im is not None else b_re))
    if product == "real_a":
        p.product = abi.QB_CPROD_REAL_A
        p.bcm = abi.bcm_types(resolve_rc(True, a_re, b_re, b_im, tf_tags))
    elif product == "real_b":
        p.product = abi.QB_CPROD_REAL_B
        p.bcm = abi.bcm_types(resolve_rc(False, b_re, a_re, a_im, tf_tags))
    elif product == "basic":
        p.product = abi.QB_CPROD_BASIC
        p.bcm = abi.bcm_types(resolve_bcm(a_re, a_im, b_re, b_im, tf_tags))
    else:
        p.tf = abi.tfc_types(resolve_tf(a_re, a_im, b_re, b_im, tf_tags))
    p.add_re, p.add_im = abi.levels(add_re), abi.levels(add_im if add_im is not None else add_re)
    p.c_re, p.c_im = abi.desc(c_re), abi.desc(c_im)
    p.M, p.N, p.K = M, N, K
    p.trans_a, p.trans_b, p.tree_form = int(trans_a), int(trans_b), form
    p.stride_a, p.stride_b, p.stride_c = M * K, K * N, M * N
    return p


def _cdesc(re, im):
    """Complex buffers are AoS pairs; both parts use the wider of the two natural words (qb200.h)."""
    wb = max(re.word_bytes or abi.qtype_of(re).word_bytes, im.word_bytes or abi.qtype_of(im).word_bytes)
    d = abi.qb_qdesc()
    d.int_bits, d.frac_bits, d.is_signed, d.qu_mode, d.of_mode, d.word_bytes = re.int_bits, re.frac_bits, re.is_signed, re.qu_mode, re.of_mode, wb
    return d


def Qcgemul(Cbuf, A, B, plan: abi.qb_cgemul_plan, batch: int = 1, flags: int = 0):
    lib = abi.load()
    real_a, real_b = plan.product == abi.QB_CPROD_REAL_A, plan.product == abi.QB_CPROD_REAL_B
    _check_operand(A, plan.a_re if real_a else _cdesc(plan.a_re, plan.a_im), 1 if real_a else 2, batch, plan.stride_a, plan.M * plan.K, "A")
    _check_operand(B, plan.b_re if real_b else _cdesc(plan.b_re, plan.b_im), 1 if real_b else 2, batch, plan.stride_b, plan.K * plan.N, "B")
    _check_operand(Cbuf, _cdesc(plan.c_re, plan.c_im), 2, batch, plan.stride_c, plan.M * plan.N, "C")
    ex = _exec_for(Cbuf, A, B, flags=flags)
    abi.check(lib.qb200_cgemul_batched(C.byref(plan), _ptr(A), _ptr(B), _ptr(Cbuf), batch, C.byref(ex)))
    return Cbuf


def tfcmul(x, y, x_re: QType, x_im: QType, y_re: QType, y_im: QType, tf_tags=None, batch: int = 1, product: str = "tf", flags: int = 0):
    """Element-wise Qmul<TFComplexMul<tags...>>(x, y) (or BasicComplexMul with product="basic") on complex vectors (AoS pairs);
    product="real_a" / "real_b": x resp. y is a REAL vector (one word per element) scaling the complex one."""
    n = (y if product == "real_a" else x)
    n = (n.numel() if _is_torch(n) else n.size) // 2
    p = abi.qb_tfcmul_plan()
    p.x_re, p.x_im, p.y_re, p.y_im = (abi.desc(t) for t in (x_re, x_im if x_im is not None else x_re, y_re, y_im if y_im is not None else y_re))
    if product == "real_a":
        r = resolve_rc(True, x_re, y_re, y_im, tf_tags)
        p.bcm = abi.bcm_types(r); p.product = abi.QB_CPROD_REAL_A
        o_re, o_im = r["ac"], r["ad"]
    elif product == "real_b":
        r = resolve_rc(False, y_re, x_re, x_im, tf_tags)
        p.bcm = abi.bcm_types(r); p.product = abi.QB_CPROD_REAL_B
        o_re, o_im = r["ac"], r["bc"]
    elif product == "basic":
        r = resolve_bcm(x_re, x_im, y_re, y_im, tf_tags)
        p.bcm = abi.bcm_types(r); p.product = abi.QB_CPROD_BASIC
        o_re, o_im = r["acbd"], r["adbc"]
    else:
        r = resolve_tf(x_re, x_im, y_re, y_im, tf_tags)
        p.tf = abi.tfc_types(r)
        o_re, o_im = r["AB"], r["BC"]
    p.len = n // batch
    out = _new_like(x, 2 * n, max(o_re.word_bytes, o_im.word_bytes))
    ex = _exec_for(x, y, flags=flags)
    abi.check(abi.load().qb200_tfcmul_batched(C.byref(p), _ptr(x), _ptr(y), _ptr(out), batch, C.byref(ex)))
    return out, (o_re, o_im)


def make_ew_plan(op: str, ta: QType, tb: QType, n: int, to=None) -> abi.qb_ew_plan:
    """op in {"mul", "add", "sub"}; `to` = merger arguments (full QType, dict of partial tags, or None for the defaults)."""
    p = abi.qb_ew_plan()
    rt = to if isinstance(to, QType) else (mul_merge if op == "mul" else add_merge)(ta, tb, to)
    p.a, p.b, p.to, p.len = abi.desc(ta), abi.desc(tb), abi.desc(rt), n
    p.op = {"mul": abi.QB_EW_MUL, "add": abi.QB_EW_ADD, "sub": abi.QB_EW_SUB}[op]
    return p


def elementwise(op: str, a, b, ta: QType, tb: QType, to=None, flags: int = 0):
    """Element-wise Qmul / Qadd / Qsub<to>(a, b) of two real tensors on the GPU; returns (out, result type)."""
    n = a.numel() if _is_torch(a) else a.size
    p = make_ew_plan(op, ta, tb, n, to)
    rt = abi.qtype_of(p.to)
    out = _new_like(a, n, rt.word_bytes)
    ex = _exec_for(a, b, flags=flags)
    abi.check(abi.load().qb200_elementwise_batched(C.byref(p), _ptr(a), _ptr(b), _ptr(out), 1, C.byref(ex)))
    return out, rt


def convert(x, src: QType, dst: QType, flags: int = 0):
    n = x.numel() if _is_torch(x) else x.size
    p = abi.qb_convert_plan()
    p.from_, p.to, p.len = abi.desc(src), abi.desc(dst), n
    out = _new_like(x, n, dst.word_bytes)
    ex = _exec_for(x, flags=flags)
    abi.check(abi.load().qb200_convert_batched(C.byref(p), _ptr(x), _ptr(out), 1, C.byref(ex)))
    return out


def make_poly_plan(x_t: QType, segments, n: int = 0, approx: Optional[bool] = None) -> abi.qb_poly_plan:
    """ANUS::Qpoly / Qapprox (reference QuBLAS.h:4829-4897).  segments = [(breakpoint, [(QType, raw), ...]), ...] with the
    coefficients in the reference's order a0, a1, ...; one segment with approx=False is Qpoly<a0, a1, ...>(x)."""
    p = abi.qb_poly_plan()
    p.x, p.len, p.nseg = abi.desc(x_t), n, len(segments)
    p.approx = int(len(segments) > 1 if approx is None else approx)
    if not 1 <= len(segments) <= abi.QB_POLY_MAX_SEG:
        raise ValueError("1 .. %d segments" % abi.QB_POLY_MAX_SEG)
    for k, (bp, coefs) in enumerate(segments):
        if not 1 <= len(coefs) <= abi.QB_POLY_MAX_COEF:
            raise ValueError("1 .. %d coefficients per segment" % abi.QB_POLY_MAX_COEF)
        p.seg[k].breakpoint, p.seg[k].ncoef = float(bp), len(coefs)
        for i, (t, raw) in enumerate(coefs):
            p.seg[k].coef[i].type, p.seg[k].coef[i].raw = abi.desc(t), int(raw)
    return p


def plan_info_poly(plan):
    info = abi.qb_plan_info()
    abi.check(abi.load().qb200_poly_plan_info(C.byref(plan), C.byref(info)))
    return info


def Qpoly(x, x_t: QType, segments, approx: Optional[bool] = None):
    """Polynomial / piecewise polynomial of every element of x on the GPU; returns (out, result type)."""
    n = x.numel() if _is_torch(x) else x.size
    p = make_poly_plan(x_t, segments, n, approx)
    rt = abi.qtype_of(plan_info_poly(p).out_type)
    _check_words(x, x_t.word_bytes, "x")
    out = _new_like(x, n, rt.word_bytes)
    ex = _exec_for(x)
    abi.check(abi.load().qb200_poly_batched(C.byref(p), _ptr(x), _ptr(out), 1, C.byref(ex)))
    return out, rt


def Qgemul_poly(Cbuf, A, B, plan: abi.qb_gemul_plan, poly: abi.qb_poly_plan, batch: int = 1, flags: int = 0):
    """Qgemul with the polynomial applied to every result element before it is stored (plan.c must be poly.x)."""
    ex = _exec_for(A, B, Cbuf, flags=flags)
    abi.check(abi.load().qb200_gemul_poly_batched(C.byref(plan), C.byref(poly), _ptr(A), _ptr(B), _ptr(Cbuf), batch, C.byref(ex)))
    return Cbuf


def Qgemul_prologue(Cbuf, A, B, plan: abi.qb_gemul_plan, poly_a=None, poly_b=None, batch: int = 1, flags: int = 0):
    """Qgemul(C, poly_a(A), poly_b(B)): the polynomial(s) applied to the operand elements as they are staged (one launch).
    A polynomial's result type must be the plan's operand type; None = that operand is used as it is."""
    ex = _exec_for(A, B, Cbuf, flags=flags)
    abi.check(abi.load().qb200_gemul_prologue_batched(C.byref(plan), C.byref(poly_a) if poly_a is not None else None,
                                                      C.byref(poly_b) if poly_b is not None else None, _ptr(A), _ptr(B), _ptr(Cbuf), batch, C.byref(ex)))
    return Cbuf


def Qtable(x, x_t: QType, out_t: QType, func):
    """ANUS::Qtable<func>(x) (reference readme.md:66-78; not in the header): out = out_t(func(x.toDouble())), through a
    table over every raw value of x_t built here with the library's own double -> Qu quantizer."""
    lo, hi = x_t.raw_min, x_t.raw_max
    raws = np.arange(lo, hi + 1, dtype=np.float64)
    with np.errstate(all="ignore"):
        vals = np.asarray(func(raws * 2.0 ** (-x_t.f)), dtype=np.float64)
    vals = np.where(np.isfinite(vals), vals, 0.0)
    if _is_torch(x):
        table = quantize(torch.from_numpy(vals).to(x.device), out_t)
    else:
        table = quantize(vals, out_t)
    n = x.numel() if _is_torch(x) else x.size
    out = _new_like(x, n, out_t.word_bytes)
    dx, do = abi.desc(x_t), abi.desc(out_t)
    ex = _exec_for(x)
    abi.check(abi.load().qb200_table_batched(C.byref(dx), C.byref(do), _ptr(table), _ptr(x), _ptr(out), n, C.byref(ex)))
    return out


def quantize(x, t: QType):
    """double -> Qu<t> raw words on the GPU: exact value, t's QuMode, t's OfMode (Qu_s(double), QuBLAS.h:2387-2393)."""
    n = x.numel() if _is_torch(x) else x.size
    out = _new_like(x, n, t.word_bytes)
    d = abi.desc(t)
    ex = _exec_for(x)
    abi.check(abi.load().qb200_quantize_f64(C.byref(d), _ptr(x), _ptr(out), n, C.byref(ex)))
    return out


def dequantize(w, t: QType):
    """raw words -> double (toDouble, QuBLAS.h:2413-2416)."""
    n = w.numel() if _is_torch(w) else w.size
    out = torch.empty(n, dtype=torch.float64, device=w.device) if _is_torch(w) else np.empty(n, dtype=np.float64)
    d = abi.desc(t)
    ex = _exec_for(w)
    abi.check(abi.load().qb200_dequantize_f64(C.byref(d), _ptr(w), _ptr(out), n, C.byref(ex)))
    return out


def error_stats(w, ref, t: QType):
    """SQNR (dB) and max abs error of a fixed-point buffer against a float64 reference, computed on the device."""
    import math
    n = w.numel() if _is_torch(w) else w.size
    st = abi.qb_err_stats()
    d = abi.desc(t)
    ex = _exec_for(w, ref)
    abi.check(abi.load().qb200_error_stats(C.byref(d), _ptr(w), _ptr(ref), n, C.byref(st), C.byref(ex)))
    sqnr = float("inf") if st.sum_sq_err == 0 else 10.0 * math.log10(st.sum_sq_ref / st.sum_sq_err)
    return {"sqnr_db": sqnr, "max_abs_err": st.max_abs_err, "sum_sq_err": st.sum_sq_err, "sum_sq_ref": st.sum_sq_ref, "n": st.n}


def gemul_f64(A, B, plan: abi.qb_gemul_plan, batch: int = 1):
    """float64 GEMM of the dequantized operands of a real Qgemul plan (batch x M x N doubles, column-major)."""
    n = batch * plan.M * plan.N
    if _is_torch(A):
        import torch
        out = torch.empty(n, dtype=torch.float64, device=A.device)
    else:
        out = np.empty(n, dtype=np.float64)
    ex = _exec_for(A, B, out)
    abi.check(abi.load().qb200_gemul_f64_batched(C.byref(plan), _ptr(A), _ptr(B), _ptr(out), batch, C.byref(ex)))
    return out


def gemul_error(Cbuf, A, B, plan: abi.qb_gemul_plan, batch: int = 1):
    """Error statistics of quantized results against the float64 GEMM of the dequantized operands, in one pass on the device
    (qb200_gemul_error_batched): nothing but the sums comes back."""
    import math
    st = abi.qb_err_stats()
    ex = _exec_for(Cbuf, A, B)
    abi.check(abi.load().qb200_gemul_error_batched(C.byref(plan), _ptr(A), _ptr(B), _ptr(Cbuf), batch, C.byref(st), C.byref(ex)))
    sqnr = float("inf") if st.sum_sq_err == 0 else 10.0 * math.log10(st.sum_sq_ref / st.sum_sq_err)
    return {"sqnr_db": sqnr, "max_abs_err": st.max_abs_err, "sum_sq_err": st.sum_sq_err, "sum_sq_ref": st.sum_sq_ref, "n": st.n}


def synth_fill(out, lo: int, hi: int, seed: int, stream_id: int, first_index: int = 0):
    """Counter-based synthetic raw words, uniform over [lo, hi]."""
    n = out.numel() if _is_torch(out) else out.size
    wb = out.element_size() if _is_torch(out) else out.dtype.itemsize
    ex = _exec_for(out)
    abi.check(abi.load().qb200_synth_fill(_ptr(out), n, wb, lo, hi, seed, stream_id, first_index, C.byref(ex)))
    return out


def shard_range(total: int, world: int, rank: int):
    f, c = C.c_uint64(), C.c_uint64()
    abi.check(abi.load().qb200_shard_range(total, world, rank, C.byref(f), C.byref(c)))
    return f.value, c.value
