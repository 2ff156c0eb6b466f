"""The compiled reference (oracle/_ref/libqref_bench.so = the unmodified pikapuma/QuBLAS header, built by oracle/Makefile)
against the C oracle at the BASELINE types and sizes:

  config 3: Qmul<mul_t> + variadic Qreduce<list> over K = 64 + converting assignment, full 64^3 instances incl. min x min rows
  config 4: Qmul<TFComplexMul<...>> + two real vector-form reductions per row, K = 256, half-range and full-range inputs
  config 2: three chained vector-form reductions 512 * 512 * 64 = 2^24 with level-shifted TypeLists (C2a, C2b; three modes),
            including all 2^15 level-9 partials (a wrong low level is often masked in the root)

CPU only.  Skipped where oracle/_ref is absent (it is built in the container that has /root/reference and travels)."""
import ctypes
import os

import numpy as np
import pytest

from oracle import oracle
from qublas_b200 import ops
from qublas_b200.qtypes import QType, RND_CONV, RND_POS_INF, SAT_TCPL, TREE_VARIADIC, TRN_TCPL

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "libqref_bench.so")
pytestmark = pytest.mark.skipif(not os.path.exists(SO), reason="oracle/_ref not built (no reference tree on this host)")
vp = ctypes.c_void_p


@pytest.fixture(scope="module")
def ref():
    lib = ctypes.CDLL(SO)
    lib.qref_c3_gemul.argtypes = [vp, vp, vp, ctypes.c_int]
    lib.qref_c4_matvec.argtypes = [vp, vp, vp, ctypes.c_int]
    lib.qref_c2_chain.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp]
    return lib


def _p(a):
    return a.ctypes.data_as(vp)


def test_c3_reference_equals_oracle_on_full_instances(ref):
    a = QType(7, 8)
    mul = QType(14, 8, True, RND_CONV, SAT_TCPL)
    lv = [QType(15 + l, 8 - l, True, RND_POS_INF, SAT_TCPL) for l in range(6)]
    c = QType(17, 4, True, RND_POS_INF, SAT_TCPL)
    plan = ops.make_gemul_plan(a, a, c, 64, 64, 64, mul=mul, add=lv)
    n = 72
    A = oracle.synth_fill(n * 4096, 4, a.raw_min, a.raw_max, 1, 0)
    B = oracle.synth_fill(n * 4096, 4, a.raw_min, a.raw_max, 1, 1)
    # saturation corners: an instance of min x min only, rows / columns of min inside random instances, and max x max
    A[:4096] = a.raw_min; B[:4096] = a.raw_min
    A[4096:8192].reshape(64, 64)[:, 5] = a.raw_min; B[4096:8192].reshape(64, 64)[9, :] = a.raw_min
    A[8192:12288] = a.raw_max; B[8192:12288] = a.raw_max
    A[12288:16384] = a.raw_min; B[12288:16384] = a.raw_max
    got = np.zeros(n * 4096, dtype=np.int32)
    ref.qref_c3_gemul(_p(A), _p(B), _p(got), n)
    want = oracle.gemul_batched(plan, A, B, n)
    assert np.array_equal(got, want)
    assert len(np.unique(got)) > 1000 and got[:4096].max() == c.raw_max       # the min x min instance saturates C


@pytest.mark.parametrize("half_range", [True, False])
def test_c4_reference_equals_oracle(ref, half_range):
    t = QType(3, 12); P = QType(6, 12); S = QType(7, 12)
    tags = {"abT": QType(4, 12), "cdT": QType(4, 11), "abcT": P, "cdbT": P, "badT": P, "ABT": S, "BCT": S}
    lv = [QType(8 + l, 12 - l, True, RND_CONV, SAT_TCPL) for l in range(8)]
    plan = ops.make_cgemul_plan(t, t, t, t, lv[-1], lv[-1], 256, 1, 256, tf_tags=tags, add_re=lv, form=TREE_VARIADIC)
    n = 3
    lo, hi = (-(1 << 14), (1 << 14) - 1) if half_range else (t.raw_min, t.raw_max)     # full range: b - a saturates
    H = oracle.synth_fill(n * 256 * 256 * 2, 4, lo, hi, 5, 0)
    x = oracle.synth_fill(n * 256 * 2, 4, lo, hi, 5, 1)
    got = np.zeros(n * 256 * 2, dtype=np.int32)
    ref.qref_c4_matvec(_p(H), _p(x), _p(got), n)
    want = oracle.cgemul_batched(plan, H, x, n)
    assert np.array_equal(got, want)
    assert len(np.unique(got)) > 200


@pytest.mark.parametrize("variant,mode", [(1, RND_POS_INF), (1, RND_CONV), (1, TRN_TCPL), (0, RND_POS_INF)])
def test_c2_chained_reference_equals_oracle_at_three_depths(ref, variant, mode):
    n = 1 << 24
    elem = QType(8, 8)
    lv = [QType(9 + l, 8 - l, True, mode, SAT_TCPL) for l in range(8)]
    if variant == 1:
        lv += [QType(17 + (l - 6) // 2, 0, True, mode, SAT_TCPL) for l in range(8, 24)]
    x = oracle.synth_fill(n, 4, elem.raw_min, elem.raw_max, 1, 3 + variant)
    root = np.zeros(1, dtype=np.int32); l9 = np.zeros(1 << 15, dtype=np.int32)
    assert ref.qref_c2_chain(_p(x), variant, mode, _p(root), _p(l9)) == 0
    plan = ops.make_reduce_plan(elem, lv, n)
    want, parts = oracle.reduce_batched(plan, x, 1, partial_level=9)
    assert int(want[0]) == int(root[0])
    assert np.array_equal(parts.astype(np.int64), l9.astype(np.int64))
    _, p18 = oracle.reduce_batched(plan, x, 1, partial_level=18)
    # level 18 through the header: reduce the level-9 partials with the shifted list (checked against the oracle's own level 18)
    plan9 = ops.make_reduce_plan(lv[8], lv[9:], 1 << 15) if variant == 1 else ops.make_reduce_plan(lv[7], [lv[7]], 1 << 15)
    want9, p9 = oracle.reduce_batched(plan9, l9, 1, partial_level=9)
    assert int(want9[0]) == int(root[0]) and np.array_equal(p9, p18)
