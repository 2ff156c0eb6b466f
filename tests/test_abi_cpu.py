"""CPU-side checks of the C ABI: the library loads, exports every symbol include/qb200.h declares,
validates plans without a device, fences the parity domain, and refuses to compute without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from qublas_b200 import abi, ops
from qublas_b200.qtypes import QType, RND_CONV, RND_POS_INF, SAT_TCPL, TREE_VARIADIC, TREE_VECTOR, WRP_TCPL

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "qb200.h")).read()
    declared = set(re.findall(r"\b(qb200_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(abi.ABI_SYMBOLS), declared ^ set(abi.ABI_SYMBOLS)
    lib = abi.load()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.qb200_abi_version() == 2


def test_struct_layout_matches_header(tmp_path):
    """ctypes mirrors == the C compiler's view of include/qb200.h."""
    import subprocess
    names = ["qb_qdesc", "qb_levels", "qb_reduce_plan", "qb_gemul_plan", "qb_tfc_types", "qb_bcm_types", "qb_cgemul_plan", "qb_err_stats",
             "qb_tfcmul_plan", "qb_ew_plan", "qb_convert_plan", "qb_plan_info", "qb_exec", "qb_shard"]
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "qb200.h"\nint main(void){' +
                   "".join('printf("%%zu\\n", sizeof(%s));' % n for n in names) + "return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    for n, sz in zip(names, sizes):
        assert C.sizeof(getattr(abi, n)) == sz, n


def test_plan_info_reduce_c2():
    lv = [QType(9 + l, 8 - l, True, RND_POS_INF, SAT_TCPL) for l in range(8)]
    info = ops.plan_info_reduce(ops.make_reduce_plan(QType(8, 8), lv, 1 << 24))
    assert info.arith_bits == 32 and info.kernel == b"k_reduce_ct<C2a,POS_INF>" and info.specialized == 1
    other = ops.plan_info_reduce(ops.make_reduce_plan(QType(8, 8), lv, (1 << 24) - 1))
    assert other.kernel == b"k_reduce_warp" and other.specialized == 0
    assert abi.qtype_of(info.out_type) == lv[-1] and info.out_word_bytes == 4
    assert not info.lossless


def test_plan_info_gemul_c3_interval_analysis():
    a = QType(7, 8)
    mul = QType(14, 8, True, RND_CONV, SAT_TCPL)
    lv = [QType(15 + l, 8 - l, True, RND_POS_INF, SAT_TCPL) for l in range(6)]
    c = QType(17, 4, True, RND_POS_INF, SAT_TCPL)
    info = ops.plan_info_gemul(ops.make_gemul_plan(a, a, c, 64, 64, 64, mul=mul, add=lv))
    assert info.arith_bits == 32 and info.max_inter_bits == 32 and not info.lossless and not info.tensor_core


def test_plan_info_gemul_c5_lossless_predicate():
    a = QType(7, 0)
    acc = QType(30, 0)
    p = ops.make_gemul_plan(a, a, acc, 256, 256, 256, mul=QType(15, 0), add=[acc])
    info = ops.plan_info_gemul(p)
    assert info.lossless and info.tensor_core and info.arith_bits == 32
    # FullPrec product type 14.0 loses min*min -> a clamp is live -> not lossless
    p2 = ops.make_gemul_plan(a, a, acc, 256, 256, 256, mul={"full": True}, add=[acc])
    info2 = ops.plan_info_gemul(p2)
    assert not info2.lossless and not info2.tensor_core


@pytest.mark.parametrize("shift,mode,ok", [(32, RND_POS_INF, False), (31, RND_CONV, False), (31, RND_POS_INF, True),
                                           (32, 5, True), (33, RND_CONV, True)])
def test_parity_domain_fence_shifts(shift, mode, ok):
    """SURVEY.md 2.3 defects 1-2: RND shifts of 32/64 and CONV shifts of 31/63 are outside the domain."""
    elem = QType(10, 40)
    lv = [QType(12, 40 - shift, True, mode, SAT_TCPL)]
    plan = ops.make_reduce_plan(elem, lv, 2)
    info = abi.qb_plan_info()
    rc = abi.load().qb200_reduce_plan_info(C.byref(plan), C.byref(info))
    assert (rc == 0) == ok, abi.load().qb200_last_error()
    if not ok:
        assert rc == abi.QB_EDOMAIN


def test_parity_domain_fence_width_and_wrap():
    lib = abi.load()
    info = abi.qb_plan_info()
    # two-word slice: the product of two 40-bit types is an 80-bit ArbiInt in the reference; after the shift by 20
    # it fits 63 bits -> inside the parity domain (tests/golden/ref_wide.json), evaluated in __int128
    p = ops.make_gemul_plan(QType(19, 20), QType(19, 20), QType(20, 20), 2, 2, 2, mul=QType(20, 20))
    assert lib.qb200_gemul_plan_info(C.byref(p), C.byref(info)) == 0 and info.arith_bits == 64 and info.max_inter_bits > 64
    # ... but not when the value at the overflow compare can reach [2^63, 2^64) (reference defect 4)
    p = ops.make_gemul_plan(QType(19, 20), QType(19, 20), QType(20, 36), 2, 2, 2, mul=QType(20, 36))
    assert lib.qb200_gemul_plan_info(C.byref(p), C.byref(info)) == abi.QB_EDOMAIN
    # ... nor with RND::CONV on the wide intermediate (defect 3), nor beyond 128 bits
    p = ops.make_gemul_plan(QType(19, 20), QType(19, 20), QType(20, 20), 2, 2, 2, mul=QType(20, 20, True, RND_CONV, SAT_TCPL))
    assert lib.qb200_gemul_plan_info(C.byref(p), C.byref(info)) == abi.QB_EDOMAIN
    # unsigned WRP::TCPL with int+frac == 32 (reference mask is a no-op)
    r = ops.make_reduce_plan(QType(20, 14), [QType(20, 12, False, 5, WRP_TCPL)], 4)
    assert lib.qb200_reduce_plan_info(C.byref(r), C.byref(info)) == abi.QB_EDOMAIN
    # malformed
    bad = ops.make_reduce_plan(QType(8, 8), None, 0)
    assert lib.qb200_reduce_plan_info(C.byref(bad), C.byref(info)) == abi.QB_EINVAL


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 1000000, 8192):
        for world in (1, 2, 3, 8):
            got = [ops.shard_range(total, world, r) for r in range(world)]
            assert got[0][0] == 0 and sum(c for _, c in got) == total
            for (f0, c0), (f1, _) in zip(got, got[1:]):
                assert f0 + c0 == f1


def test_compute_fails_loudly_without_gpu():
    lib = abi.load()
    if lib.qb200_device_count() > 0:
        pytest.skip("a GPU is present")
    x = np.arange(8, dtype=np.int32)
    with pytest.raises(abi.QbError) as e:
        ops.Qreduce(x, QType(8, 8), [QType(9, 8)])
    assert e.value.code == abi.QB_ENODEVICE and "no CPU fallback" in str(e.value)


def test_jit_specialisation_compiles_without_a_device():
    """Run-time specialisation: codegen + NVRTC produce an sm_100a cubin for arbitrary plans (no GPU needed)."""
    import ctypes
    lib = abi.load()
    log = ctypes.create_string_buffer(8192)
    rng = np.random.default_rng(0)
    for (M, N, K) in [(32, 16, 32), (64, 64, 64), (16, 32, 8)]:
        a = QType(int(rng.integers(3, 8)), int(rng.integers(2, 8)), True, int(rng.integers(0, 7)), int(rng.integers(0, 4)))
        mul = QType(a.i * 2, a.f, True, int(rng.integers(0, 7)), int(rng.integers(0, 4)))
        lv = [QType(mul.i + l, max(mul.f - l, 0), True, int(rng.integers(0, 7)), int(rng.integers(0, 4))) for l in range(4)]
        plan = ops.make_gemul_plan(a, a, lv[-1], M, N, K, mul=mul, add=lv)
        rc = lib.qb200_jit_check_gemul(ctypes.byref(plan), log, 8192)
        assert rc == 0, log.value.decode()
    # shapes the register-tile template cannot take (ragged M/N, K not a power of two, transposes, both tree forms, the
    # 64-bit register class) compile through the thread-per-output template
    from qublas_b200.qtypes import TREE_VARIADIC, TREE_VECTOR
    for (M, N, K, form, ta, tb, wide) in [(5, 7, 12, TREE_VECTOR, False, False, False), (48, 48, 48, TREE_VARIADIC, True, False, False),
                                          (64, 1, 100, TREE_VECTOR, False, True, False), (2, 2, 1, TREE_VECTOR, False, False, False),
                                          (8, 8, 12, TREE_VARIADIC, False, False, True)]:
        a = QType(20, 10) if wide else QType(7, 8)
        mul = QType(40, 12, True, 4, 0) if wide else QType(14, 8, True, 4, 0)
        lv = [QType(mul.i + 1 + l, max(mul.f - l, 2), True, int(rng.integers(0, 7)), int(rng.integers(0, 4))) for l in range(3)]
        plan = ops.make_gemul_plan(a, a, lv[-1], M, N, K, mul=mul, add=lv, form=form, trans_a=ta, trans_b=tb)
        rc = lib.qb200_jit_check_gemul(ctypes.byref(plan), log, 8192)
        assert rc == 0, (M, N, K, log.value.decode())


def test_register_tile_variants_compile_without_a_device():
    """The register-tile template in the forms its compile-time predicates select (qb_fast_gemul.cuh): substituted operands
    (fp_subst) with the negated deferred-shift tree (all levels RND::POS_INF), with group constants (mixed level modes), without
    the deferred-shift form (wide levels), at K = 128, with packed int16 words, with groups of small instances per CTA, with
    transposed operands; and the per-instance clamp elision where the substitution does not apply (TRN::TCPL product)."""
    import ctypes
    from qublas_b200.qtypes import RND_CONV, RND_NEG_INF, RND_POS_INF, SAT_TCPL, TRN_TCPL
    lib = abi.load()
    log = ctypes.create_string_buffer(16384)
    a = QType(7, 8); c = QType(17, 4, True, RND_POS_INF, SAT_TCPL)
    conv = QType(14, 8, True, RND_CONV, SAT_TCPL); trn = QType(14, 8, True, TRN_TCPL, SAT_TCPL)

    def halving(K, modes):
        return [QType(15, 8, True, RND_POS_INF, SAT_TCPL)] + [QType(15 + l, 8 - l, True, modes[(l - 1) % len(modes)], SAT_TCPL) for l in range(1, K.bit_length() - 1)]

    cases = [
        (64, 64, 64, conv, halving(64, [RND_POS_INF]), {}),
        (64, 64, 64, conv, halving(64, [RND_POS_INF, RND_NEG_INF, TRN_TCPL]), {}),
        (64, 64, 64, conv, [QType(16 + l, 8, True, RND_POS_INF, SAT_TCPL) for l in range(6)], {}),
        (64, 64, 128, conv, halving(128, [RND_POS_INF]), {}),
        (64, 64, 64, conv, halving(64, [RND_POS_INF]), {"word_bytes": (2, 2, 0)}),
        (16, 16, 16, conv, halving(16, [RND_POS_INF]), {}),
        (64, 32, 64, conv, halving(64, [RND_POS_INF]), {"trans_a": True, "trans_b": True}),
        (64, 64, 64, trn, halving(64, [RND_POS_INF]), {}),
    ]
    for (M, N, K, mul, lv, kw) in cases:
        plan = ops.make_gemul_plan(a, a, c, M, N, K, mul=mul, add=lv, **kw)
        rc = lib.qb200_jit_check_gemul(ctypes.byref(plan), log, 16384)
        assert rc == 0, (M, N, K, kw, log.value.decode()[:1500])


def test_every_jit_template_compiles_without_a_device():
    """qb200_jit_check: the complex thread-per-output template (all four product kinds), the mat-vec rows template, both
    Qreduce templates, and the element-wise templates (complex products, real ops, convert) go through codegen + NVRTC
    on a machine without a GPU."""
    import ctypes
    from qublas_b200.qtypes import TREE_VARIADIC, TREE_VECTOR
    lib = abi.load()
    log = ctypes.create_string_buffer(16384)

    def check(what, plan):
        rc = lib.qb200_jit_check(what, ctypes.cast(ctypes.byref(plan), ctypes.c_void_p), log, 16384)
        assert rc == 0, (what, rc, log.value.decode()[:2000])

    t = QType(3, 12); P = QType(6, 12); S = QType(7, 12)
    tf = {"abT": QType(4, 12), "cdT": QType(4, 11), "abcT": P, "cdbT": P, "badT": P, "ABT": S, "BCT": S}
    bc = {"acT": P, "bdT": QType(6, 11, True, 4, 1), "adT": P, "bcT": P, "acbdT": S, "adbcT": QType(7, 10, True, 0, 2)}
    rc_tags = {"realT": P, "imagT": QType(6, 10, True, 2, 0)}
    lv = [QType(8 + l, 12 - l, True, 4, 0) for l in range(4)]
    out = QType(12, 8)
    # complex Qgemul: mat-vec rows template, then the thread template for every product kind
    check(1, ops.make_cgemul_plan(t, t, t, t, out, out, 64, 1, 64, tf_tags=tf, add_re=lv))
    for kind, tags, args in (("tf", tf, (t, t, t, t)), ("basic", bc, (t, t, t, t)), ("real_a", rc_tags, (t, None, t, t)), ("real_b", rc_tags, (t, t, t, None))):
        check(1, ops.make_cgemul_plan(*args, out, out, 5, 3, 12, tf_tags=tags, add_re=lv, form=TREE_VECTOR, product=kind, trans_a=True))
    # Qreduce: power-of-two two-pass template, any-length template, variadic power of two routed to the vector kernels
    elem = QType(8, 8); rl = [QType(9 + l, 8 - l, True, 0, 0) for l in range(8)]
    # ... and the variadic form of lengths that are not a power of two (whole 256-node chunks per pass + finishing kernel)
    for n, form in ((1 << 16, TREE_VECTOR), (100000, TREE_VECTOR), (64, TREE_VECTOR), (4097, TREE_VECTOR), (1 << 12, TREE_VARIADIC),
                    (100000, TREE_VARIADIC), (3 << 20, TREE_VARIADIC), (513, TREE_VARIADIC)):
        check(2, ops.make_reduce_plan(elem, rl, n, form=form))
    # element-wise: complex products, real ops, converting assignment
    for kind, tags, args in (("tf", tf, (t, t, t, t)), ("basic", bc, (t, t, t, t)), ("real_a", rc_tags, (t, None, t, t))):
        p = abi.qb_tfcmul_plan()
        p.x_re, p.x_im, p.y_re, p.y_im = (abi.desc(q if q is not None else t) for q in args)
        if kind == "tf":
            from qublas_b200.qtypes import resolve_tf
            p.tf = abi.tfc_types(resolve_tf(t, t, t, t, tags))
        elif kind == "basic":
            from qublas_b200.qtypes import resolve_bcm
            p.bcm = abi.bcm_types(resolve_bcm(t, t, t, t, tags)); p.product = abi.QB_CPROD_BASIC
        else:
            from qublas_b200.qtypes import resolve_rc
            p.bcm = abi.bcm_types(resolve_rc(True, t, t, t, tags)); p.product = abi.QB_CPROD_REAL_A
        p.len = 100
        check(3, p)
    for op in ("mul", "add", "sub"):
        check(4, ops.make_ew_plan(op, QType(7, 8), QType(5, 6), 100, QType(10, 6, True, 4, 0)))
    cp = abi.qb_convert_plan(); cp.from_, cp.to, cp.len = abi.desc(QType(20, 10)), abi.desc(QType(36, 9, True, 3, 2)), 10
    check(5, cp)


def test_jit_disk_cache(tmp_path, monkeypatch):
    """QB200_JIT_CACHE=<dir>: compiled kernels are stored under a hash of their generated source and reused."""
    import ctypes
    import time
    monkeypatch.setenv("QB200_JIT_CACHE", str(tmp_path))
    lib = abi.load()
    log = ctypes.create_string_buffer(8192)
    plan = ops.make_gemul_plan(QType(6, 7), QType(6, 7), QType(15, 5), 32, 32, 16, mul=QType(12, 7, True, 4, 0),
                               add=[QType(13 + l, 7 - l, True, 0, 0) for l in range(4)])
    t0 = time.time(); assert lib.qb200_jit_check_gemul(ctypes.byref(plan), log, 8192) == 0; cold = time.time() - t0
    files = [f for f in os.listdir(tmp_path) if f.endswith(".cubin")]
    assert len(files) == 1 and os.path.getsize(os.path.join(tmp_path, files[0])) > 1000
    t0 = time.time(); assert lib.qb200_jit_check_gemul(ctypes.byref(plan), log, 8192) == 0; warm = time.time() - t0
    assert warm < cold and [f for f in os.listdir(tmp_path) if f.endswith(".cubin")] == files
    # the file carries a verified header (NVRTC version, source length, second hash, cubin length): a stale, foreign or
    # truncated file under the same name is recompiled and replaced instead of being loaded
    path = os.path.join(tmp_path, files[0])
    good = open(path, "rb").read()
    assert good[:8] == b"QB200JC2"
    for bad in (good[:len(good) // 2], good[:24] + bytes([good[24] ^ 1]) + good[25:], b"\x7fELF" + good[4:], good + b"x"):
        open(path, "wb").write(bad)
        assert lib.qb200_jit_check_gemul(ctypes.byref(plan), log, 8192) == 0
        assert open(path, "rb").read() == good


def test_python_wrappers_validate_buffers():
    """a wrong dtype or a short buffer is a ValueError, not an out-of-bounds device access."""
    a = QType(7, 8)
    plan = ops.make_gemul_plan(a, a, QType(17, 4), 8, 8, 8, mul=QType(14, 8), add=[QType(20, 8)])
    ok = np.zeros(2 * 64, dtype=np.int32)
    with pytest.raises(ValueError):
        ops.Qgemul(ok.copy(), ok.astype(np.int16), ok, plan, batch=2)            # int16 words under a natural (4-byte) plan
    with pytest.raises(ValueError):
        ops.Qgemul(ok.copy(), ok[:100], ok, plan, batch=2)                         # A too short for two instances
    with pytest.raises(ValueError):
        ops.Qgemul(ok[:64].copy(), ok, ok, plan, batch=2)                          # C too short
    with pytest.raises(ValueError):
        ops.Qgemul(ok.astype(np.float32), ok, ok, plan, batch=2)
    t = QType(3, 12)
    cplan = ops.make_cgemul_plan(t, t, t, t, t, t, 4, 1, 4, product="basic")
    with pytest.raises(ValueError):
        ops.Qcgemul(np.zeros(8, dtype=np.int32), np.zeros(16, dtype=np.int32), np.zeros(8, dtype=np.int32), cplan)   # A needs 32 words
    with pytest.raises(ValueError):
        ops.Qcgemul(np.zeros(8, dtype=np.int64), np.zeros(32, dtype=np.int32), np.zeros(8, dtype=np.int32), cplan)


def test_poly_plans_are_validated_without_a_gpu():
    """ANUS plans: type rules and the parity fence are host code (no device needed)."""
    from qublas_b200 import ops
    from qublas_b200.qtypes import QType
    x = QType(4, 8); k = QType(3, 6)
    info = ops.plan_info_poly(ops.make_poly_plan(x, [(0.0, [(k, 5), (k, -7)])], 16))
    assert abi.qtype_of(info.out_type) == k and info.arith_bits == 32
    info = ops.plan_info_poly(ops.make_poly_plan(x, [(0.0, [(k, 5)]), (1.0, [(k, 1), (k, 2)])], 16))
    assert abi.qtype_of(info.out_type) == x                                          # Qapprox converts back into x's type
    with pytest.raises(abi.QbError):
        ops.plan_info_poly(ops.make_poly_plan(x, [(0.0, [(k, 5)]), (1.0, [(k, 1)])], 16, approx=False))     # Qpoly takes one segment
    with pytest.raises(abi.QbError):
        ops.plan_info_poly(ops.make_poly_plan(x, [(0.0, [(k, 1 << 20)])], 16))        # coefficient outside its type


def test_poly_kernels_specialise_with_nvrtc_without_a_gpu():
    """The specialised ANUS kernel and the Qgemul kernels with the polynomial fused into their store (register-tile and
    thread-per-output templates) compile for sm_100a (NVRTC needs no device)."""
    import ctypes
    from qublas_b200.qtypes import RND_CONV, RND_POS_INF, SAT_TCPL
    lib = abi.load()
    log = ctypes.create_string_buffer(16384)
    c = QType(17, 4, True, RND_POS_INF, SAT_TCPL)
    k_t = QType(1, 14, True, RND_POS_INF, SAT_TCPL)
    segs = [(-100.0, [(k_t, 0)]), (0.0, [(k_t, 0), (k_t, 1 << 11)]), (1e9, [(k_t, 0), (k_t, 1 << 14), (k_t, -3)])]
    poly = ops.make_poly_plan(c, segs, 64, True)
    rc = lib.qb200_jit_check(6, ctypes.cast(ctypes.byref(poly), ctypes.c_void_p), log, 16384)
    assert rc == 0, log.value.decode()
    a = QType(7, 8); mul = QType(14, 8, True, RND_CONV, SAT_TCPL)

    class Pair(ctypes.Structure):
        _fields_ = [("g", abi.qb_gemul_plan), ("p", abi.qb_poly_plan)]
    for M, N, K in ((64, 64, 64), (10, 6, 12)):
        nlev = max(1, (K - 1).bit_length())
        lv = [QType(15 + l, 8 - min(l, 8), True, RND_POS_INF, SAT_TCPL) for l in range(nlev)]
        pair = Pair(ops.make_gemul_plan(a, a, c, M, N, K, mul=mul, add=lv), poly)
        rc = lib.qb200_jit_check(7, ctypes.cast(ctypes.byref(pair), ctypes.c_void_p), log, 16384)
        assert rc == 0, log.value.decode()
        # the polynomial as a prologue of both operands (its x type = its result type = the operand type)
        pro = Pair(pair.g, ops.make_poly_plan(a, segs, 64, True))
        rc = lib.qb200_jit_check(8, ctypes.cast(ctypes.byref(pro), ctypes.c_void_p), log, 16384)
        assert rc == 0, log.value.decode()
