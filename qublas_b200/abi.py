"""ctypes view of include/qb200.h: POD plan structs + loader for the CUDA library libqb200.so.

No torch types cross this boundary: pointers and sizes only.  There is no CPU fallback — if the
CUDA library is missing, loading raises.
"""
import ctypes as C
import os
from typing import Optional, Sequence

from .qtypes import QType

QB_MAX_LEVELS = 64
QB_DEVICE_ALL = -2
QB_MEM_HOST, QB_MEM_DEVICE = 0, 1
QB_F_NO_TENSOR_CORE, QB_F_NO_SPECIALIZE = 1, 2
QB_OK, QB_EINVAL, QB_EDOMAIN, QB_ENODEVICE, QB_ECUDA, QB_ENOMEM = 0, -1, -2, -3, -4, -5


class qb_qdesc(C.Structure):
    _fields_ = [("int_bits", C.c_int16), ("frac_bits", C.c_int16), ("is_signed", C.c_uint8),
                ("qu_mode", C.c_uint8), ("of_mode", C.c_uint8), ("word_bytes", C.c_uint8)]


class qb_levels(C.Structure):
    _fields_ = [("n", C.c_uint32), ("lv", qb_qdesc * QB_MAX_LEVELS)]


class qb_reduce_plan(C.Structure):
    _fields_ = [("elem", qb_qdesc), ("levels", qb_levels), ("len", C.c_uint64),
                ("tree_form", C.c_uint32), ("reserved", C.c_uint32)]


class qb_gemul_plan(C.Structure):
    _fields_ = [("a", qb_qdesc), ("b", qb_qdesc), ("mul", qb_qdesc), ("c", qb_qdesc), ("add", qb_levels),
                ("M", C.c_uint32), ("N", C.c_uint32), ("K", C.c_uint32),
                ("trans_a", C.c_uint8), ("trans_b", C.c_uint8), ("tree_form", C.c_uint8), ("reserved", C.c_uint8),
                ("stride_a", C.c_int64), ("stride_b", C.c_int64), ("stride_c", C.c_int64)]


class qb_tfc_types(C.Structure):
    _fields_ = [(k, qb_qdesc) for k in ("ab", "cd", "ba", "abc", "bad", "cdb", "AB", "BC")]


class qb_bcm_types(C.Structure):
    _fields_ = [(k, qb_qdesc) for k in ("ac", "bd", "ad", "bc", "acbd", "adbc")]


QB_CPROD_TF, QB_CPROD_BASIC, QB_CPROD_REAL_A, QB_CPROD_REAL_B = 0, 1, 2, 3


class qb_cgemul_plan(C.Structure):
    _fields_ = [("a_re", qb_qdesc), ("a_im", qb_qdesc), ("b_re", qb_qdesc), ("b_im", qb_qdesc),
                ("tf", qb_tfc_types), ("bcm", qb_bcm_types), ("add_re", qb_levels), ("add_im", qb_levels),
                ("c_re", qb_qdesc), ("c_im", qb_qdesc),
                ("M", C.c_uint32), ("N", C.c_uint32), ("K", C.c_uint32),
                ("trans_a", C.c_uint8), ("trans_b", C.c_uint8), ("tree_form", C.c_uint8), ("product", C.c_uint8),
                ("stride_a", C.c_int64), ("stride_b", C.c_int64), ("stride_c", C.c_int64)]


class qb_tfcmul_plan(C.Structure):
    _fields_ = [("x_re", qb_qdesc), ("x_im", qb_qdesc), ("y_re", qb_qdesc), ("y_im", qb_qdesc),
                ("tf", qb_tfc_types), ("bcm", qb_bcm_types), ("len", C.c_uint64), ("product", C.c_uint32), ("reserved", C.c_uint32)]


class qb_ew_plan(C.Structure):
    _fields_ = [("a", qb_qdesc), ("b", qb_qdesc), ("to", qb_qdesc), ("len", C.c_uint64), ("op", C.c_uint32), ("reserved", C.c_uint32)]


QB_EW_MUL, QB_EW_ADD, QB_EW_SUB = 0, 1, 2


class qb_convert_plan(C.Structure):
    _fields_ = [("from_", qb_qdesc), ("to", qb_qdesc), ("len", C.c_uint64)]


QB_POLY_MAX_COEF, QB_POLY_MAX_SEG = 8, 8


class qb_poly_coef(C.Structure):
    _fields_ = [("type", qb_qdesc), ("raw", C.c_int64)]


class qb_poly_segment(C.Structure):
    _fields_ = [("breakpoint", C.c_double), ("ncoef", C.c_uint32), ("reserved", C.c_uint32), ("coef", qb_poly_coef * QB_POLY_MAX_COEF)]


class qb_poly_plan(C.Structure):
    _fields_ = [("x", qb_qdesc), ("nseg", C.c_uint32), ("approx", C.c_uint32), ("seg", qb_poly_segment * QB_POLY_MAX_SEG), ("len", C.c_uint64)]


class qb_shard(C.Structure):
    _fields_ = [("device", C.c_int32), ("reserved", C.c_int32), ("A", C.c_void_p), ("B", C.c_void_p), ("C", C.c_void_p), ("batch", C.c_uint64)]


class qb_plan_info(C.Structure):
    _fields_ = [("arith_bits", C.c_int32), ("max_inter_bits", C.c_int32), ("lossless", C.c_int32),
                ("tensor_core", C.c_int32), ("specialized", C.c_int32), ("out_word_bytes", C.c_int32),
                ("out_type", qb_qdesc), ("kernel", C.c_char * 96)]


class qb_err_stats(C.Structure):
    _fields_ = [("sum_sq_err", C.c_double), ("sum_sq_ref", C.c_double), ("max_abs_err", C.c_double), ("n", C.c_uint64)]


class qb_exec(C.Structure):
    _fields_ = [("mem_kind", C.c_int32), ("device", C.c_int32), ("stream", C.c_void_p),
                ("flags", C.c_uint32), ("reserved", C.c_uint32)]


def desc(t: QType, word_bytes: int = 0) -> qb_qdesc:
    return qb_qdesc(t.i, t.f, 1 if t.s else 0, t.q, t.o, word_bytes)


def qtype_of(d: qb_qdesc) -> QType:
    return QType(d.int_bits, d.frac_bits, bool(d.is_signed), d.qu_mode, d.of_mode)


def levels(ts: Optional[Sequence[QType]]) -> qb_levels:
    L = qb_levels()
    ts = list(ts or [])
    if len(ts) > QB_MAX_LEVELS:
        raise ValueError("too many levels")
    L.n = len(ts)
    for k, t in enumerate(ts):
        L.lv[k] = desc(t)
    return L


def bcm_types(r) -> qb_bcm_types:
    t = qb_bcm_types()
    for k in ("ac", "bd", "ad", "bc", "acbd", "adbc"):
        if k in r:                      # the real (x) complex kinds fill only the products they use
            setattr(t, k, desc(r[k]))
    return t


def tfc_types(r) -> qb_tfc_types:
    t = qb_tfc_types()
    for k in ("ab", "cd", "ba", "abc", "bad", "cdb", "AB", "BC"):
        setattr(t, k, desc(r[k]))
    return t


# every symbol include/qb200.h declares (checked by tests/test_abi_symbols.py)
ABI_SYMBOLS = [
    "qb200_abi_version", "qb200_last_error", "qb200_device_count", "qb200_launch_count",
    "qb200_launch_count_reset", "qb200_reduce_plan_info", "qb200_gemul_plan_info", "qb200_cgemul_plan_info",
    "qb200_reduce_batched", "qb200_gemul_batched", "qb200_cgemul_batched", "qb200_tfcmul_batched",
    "qb200_convert_batched", "qb200_elementwise_batched", "qb200_synth_fill", "qb200_shard_range", "qb200_measure_int_peak", "qb200_jit_check_gemul", "qb200_jit_check", "qb200_quantize_f64", "qb200_dequantize_f64", "qb200_error_stats",
    "qb200_selftest_fp", "qb200_gemul_f64_batched", "qb200_gemul_error_batched", "qb200_gemul_prologue_batched", "qb200_host_alloc", "qb200_host_free", "qb200_device_alloc", "qb200_device_free", "qb200_memcpy", "qb200_gemul_sharded", "qb200_gemul_panels_gathered", "qb200_allgather",
    "qb200_poly_plan_info", "qb200_poly_batched", "qb200_gemul_poly_batched", "qb200_table_batched",
]

_LIB = None


def lib_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libqb200.so")


def load() -> C.CDLL:
    """Load libqb200.so (built in-tree by __graft_entry__.build()).  Fails loudly if absent."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise RuntimeError("qublas_b200: CUDA library %s not built — run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)" % p)
    L = C.CDLL(p)
    L.qb200_abi_version.restype = C.c_int
    L.qb200_last_error.restype = C.c_char_p
    L.qb200_device_count.restype = C.c_int
    L.qb200_launch_count.restype = C.c_uint64
    L.qb200_launch_count_reset.restype = None
    vp, u64, u32, i64 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int64
    L.qb200_reduce_plan_info.argtypes = [C.POINTER(qb_reduce_plan), C.POINTER(qb_plan_info)]
    L.qb200_gemul_plan_info.argtypes = [C.POINTER(qb_gemul_plan), C.POINTER(qb_plan_info)]
    L.qb200_cgemul_plan_info.argtypes = [C.POINTER(qb_cgemul_plan), C.POINTER(qb_plan_info)]
    L.qb200_reduce_batched.argtypes = [C.POINTER(qb_reduce_plan), vp, vp, vp, u32, u64, C.POINTER(qb_exec)]
    L.qb200_gemul_batched.argtypes = [C.POINTER(qb_gemul_plan), vp, vp, vp, u64, C.POINTER(qb_exec)]
    L.qb200_cgemul_batched.argtypes = [C.POINTER(qb_cgemul_plan), vp, vp, vp, u64, C.POINTER(qb_exec)]
    L.qb200_tfcmul_batched.argtypes = [C.POINTER(qb_tfcmul_plan), vp, vp, vp, u64, C.POINTER(qb_exec)]
    L.qb200_convert_batched.argtypes = [C.POINTER(qb_convert_plan), vp, vp, u64, C.POINTER(qb_exec)]
    L.qb200_elementwise_batched.argtypes = [C.POINTER(qb_ew_plan), vp, vp, vp, u64, C.POINTER(qb_exec)]
    L.qb200_synth_fill.argtypes = [vp, u64, u32, i64, i64, u64, u64, u64, C.POINTER(qb_exec)]
    L.qb200_shard_range.argtypes = [u64, u32, u32, C.POINTER(u64), C.POINTER(u64)]
    L.qb200_quantize_f64.argtypes = [C.POINTER(qb_qdesc), vp, vp, u64, C.POINTER(qb_exec)]
    L.qb200_dequantize_f64.argtypes = [C.POINTER(qb_qdesc), vp, vp, u64, C.POINTER(qb_exec)]
    L.qb200_error_stats.argtypes = [C.POINTER(qb_qdesc), vp, vp, u64, C.POINTER(qb_err_stats), C.POINTER(qb_exec)]
    L.qb200_jit_check_gemul.argtypes = [C.POINTER(qb_gemul_plan), C.c_char_p, u64]
    L.qb200_jit_check.argtypes = [C.c_int, C.c_void_p, C.c_char_p, u64]
    L.qb200_measure_int_peak.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(qb_exec)]
    L.qb200_gemul_f64_batched.argtypes = [C.POINTER(qb_gemul_plan), vp, vp, vp, u64, C.POINTER(qb_exec)]
    L.qb200_gemul_error_batched.argtypes = [C.POINTER(qb_gemul_plan), vp, vp, vp, u64, C.POINTER(qb_err_stats), C.POINTER(qb_exec)]
    L.qb200_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t, u32]
    L.qb200_host_free.argtypes = [vp]
    L.qb200_device_alloc.argtypes = [C.POINTER(vp), C.c_size_t, C.c_int]
    L.qb200_device_free.argtypes = [vp, C.c_int]
    L.qb200_memcpy.argtypes = [vp, vp, C.c_size_t, C.c_int]
    L.qb200_gemul_sharded.argtypes = [C.POINTER(qb_gemul_plan), C.POINTER(qb_shard), u32, C.POINTER(qb_exec)]
    L.qb200_gemul_panels_gathered.argtypes = [C.POINTER(qb_gemul_plan), C.POINTER(qb_shard), u32, C.POINTER(vp), C.POINTER(qb_exec)]
    L.qb200_allgather.argtypes = [C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_int32), u32, u64, u64, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), C.POINTER(C.c_double)]
    L.qb200_poly_plan_info.argtypes = [C.POINTER(qb_poly_plan), C.POINTER(qb_plan_info)]
    L.qb200_poly_batched.argtypes = [C.POINTER(qb_poly_plan), vp, vp, u64, C.POINTER(qb_exec)]
    L.qb200_gemul_poly_batched.argtypes = [C.POINTER(qb_gemul_plan), C.POINTER(qb_poly_plan), vp, vp, vp, u64, C.POINTER(qb_exec)]
    L.qb200_gemul_prologue_batched.argtypes = [C.POINTER(qb_gemul_plan), C.POINTER(qb_poly_plan), C.POINTER(qb_poly_plan), vp, vp, vp, u64, C.POINTER(qb_exec)]
    L.qb200_table_batched.argtypes = [C.POINTER(qb_qdesc), C.POINTER(qb_qdesc), vp, vp, vp, u64, C.POINTER(qb_exec)]
    L.qb200_selftest_fp.argtypes = [C.c_int, C.c_void_p, C.POINTER(u64), C.POINTER(u64), C.POINTER(qb_exec)]
    for name in ABI_SYMBOLS:
        getattr(L, name)
    _LIB = L
    return L


class QbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("qb200 error %d: %s" % (code, msg))
        self.code = code


def check(rc: int):
    if rc != 0:
        raise QbError(rc, (load().qb200_last_error() or b"").decode())
