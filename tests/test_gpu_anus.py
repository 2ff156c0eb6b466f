"""ANUS::Qpoly / Qapprox over tensors, as a Qgemul epilogue, and Qtable lookups on the GPU against
the reference's own outputs (tests/golden/ref_anus.txt) and against the oracle."""
import random

import numpy as np
import pytest
import torch

from oracle import oracle
from qublas_b200 import abi, ops
from qublas_b200.qtypes import QType, RND_CONV, RND_POS_INF, SAT_TCPL
from tests.helpers import anus_golden

pytestmark = pytest.mark.gpu


def test_poly_and_approx_match_reference_golden():
    x_t, exprs, xs, want = anus_golden()
    xd = torch.from_numpy(xs).cuda()
    for name, segs in exprs.items():
        out, rt = ops.Qpoly(xd, x_t, segs, approx=name in ("q", "r"))
        assert np.array_equal(out.cpu().numpy().astype(np.int64), want[name]), name
        assert rt == (x_t if name in ("q", "r") else segs[0][1][0][0])


def _rand_type(rng, wmax):
    i = rng.randint(0, wmax // 2)
    f = rng.randint(0, wmax - i)
    return QType(i, f, True, rng.randrange(7), rng.randrange(4))


@pytest.mark.parametrize("seed", range(12))
def test_random_polynomials_vs_oracle(seed):
    """Random coefficient types (all QuModes / OfModes, 32- and 64-bit register classes), 1..4 segments of degree 0..4,
    every raw value of a narrow x type or a random sample of a wide one."""
    rng = random.Random(1000 + seed)
    wide = seed % 3 == 2
    x_t = _rand_type(rng, 40 if wide else 13)
    nseg = rng.randint(1, 4)
    approx = nseg > 1 or rng.random() < 0.4
    segs = []
    bps = sorted(rng.uniform(-2.0 ** x_t.i, 2.0 ** x_t.i) for _ in range(nseg))
    for s in range(nseg):
        coefs = []
        for _ in range(rng.randint(1, 5)):
            t = _rand_type(rng, 28 if wide else 14)
            coefs.append((t, rng.randint(t.raw_min, t.raw_max)))
        segs.append((bps[s], coefs))
    if x_t.width <= 14:
        xs = np.arange(x_t.raw_min, x_t.raw_max + 1, dtype=np.int32 if x_t.word_bytes == 4 else np.int64)
    else:
        xs = np.array([rng.randint(x_t.raw_min, x_t.raw_max) for _ in range(5000)] + [x_t.raw_min, x_t.raw_max, 0, -1, 1], dtype=np.int32 if x_t.word_bytes == 4 else np.int64)
    plan = ops.make_poly_plan(x_t, segs, len(xs), approx)
    try:
        info = ops.plan_info_poly(plan)
    except abi.QbError as e:
        assert e.code == abi.QB_EDOMAIN                      # outside the parity domain: refused, not computed
        return
    want = oracle.poly_batched(plan, xs)
    xd = torch.from_numpy(xs).cuda()
    got, rt = ops.Qpoly(xd, x_t, segs, approx)                                       # specialised kernel (NVRTC)
    assert np.array_equal(got.cpu().numpy(), want), (seed, info.arith_bits)
    out2 = torch.empty_like(got)                                                      # runtime-descriptor kernel
    ex = abi.qb_exec(abi.QB_MEM_DEVICE, -1, None, abi.QB_F_NO_SPECIALIZE, 0)
    import ctypes as C
    abi.check(abi.load().qb200_poly_batched(C.byref(plan), xd.data_ptr(), out2.data_ptr(), 1, C.byref(ex)))
    assert torch.equal(got, out2)


def test_gemul_with_polynomial_epilogue_vs_oracle():
    """Qgemul whose result elements go through a piecewise polynomial before they are stored: same bits as the oracle's
    Qgemul followed by the oracle's Qapprox; device-resident and host buffers, the config-3 types at 64x64x64 and a ragged shape."""
    a = QType(7, 8); mul = QType(14, 8, True, RND_CONV, SAT_TCPL)
    c = QType(17, 4, True, RND_POS_INF, SAT_TCPL)
    k_t = QType(1, 14, True, RND_POS_INF, SAT_TCPL)
    segs = [(-100.0, [(k_t, 0)]), (0.0, [(k_t, 0), (k_t, 1 << 11)]), (1e9, [(k_t, 0), (k_t, 1 << 14), (k_t, -3)])]      # leaky-ReLU-like with a small quadratic term
    for M, N, K, batch in ((64, 64, 64, 9), (10, 6, 12, 5)):
        nlev = max(1, (K - 1).bit_length())
        lv = [QType(15 + l, 8 - min(l, 8), True, RND_POS_INF, SAT_TCPL) for l in range(nlev)]
        plan = ops.make_gemul_plan(a, a, c, M, N, K, mul=mul, add=lv)
        poly = ops.make_poly_plan(c, segs, 0, True)
        A = torch.empty(batch * M * K, dtype=torch.int32, device="cuda"); B = torch.empty(batch * K * N, dtype=torch.int32, device="cuda")
        ops.synth_fill(A, a.raw_min, a.raw_max, 5, 0); ops.synth_fill(B, a.raw_min, a.raw_max, 5, 1)
        Cd = torch.empty(batch * M * N, dtype=torch.int32, device="cuda")
        ops.Qgemul_poly(Cd, A, B, plan, poly, batch=batch)
        mid = oracle.gemul_batched(plan, A.cpu().numpy(), B.cpu().numpy(), batch)
        poly.len = mid.size
        want = oracle.poly_batched(poly, mid)
        assert np.array_equal(Cd.cpu().numpy(), want)
        Ch = np.empty(batch * M * N, dtype=np.int32)
        ops.Qgemul_poly(Ch, A.cpu().numpy(), B.cpu().numpy(), plan, poly, batch=batch)
        assert np.array_equal(Ch, want)
        C2 = torch.empty_like(Cd)                                                     # the unfused pair of runtime-descriptor kernels
        ops.Qgemul_poly(C2, A, B, plan, poly, batch=batch, flags=abi.QB_F_NO_SPECIALIZE)
        assert torch.equal(Cd, C2)


def test_fused_epilogue_is_one_launch():
    """The polynomial rides in the store of the specialised Qgemul kernel: one launch per call, not two."""
    a = QType(7, 8); mul = QType(14, 8, True, RND_CONV, SAT_TCPL)
    c = QType(17, 4, True, RND_POS_INF, SAT_TCPL)
    k_t = QType(1, 14, True, RND_POS_INF, SAT_TCPL)
    lv = [QType(15 + l, 8 - l, True, RND_POS_INF, SAT_TCPL) for l in range(6)]
    plan = ops.make_gemul_plan(a, a, c, 64, 64, 64, mul=mul, add=lv)
    poly = ops.make_poly_plan(c, [(0.0, [(k_t, 0), (k_t, 1 << 11)]), (1e9, [(k_t, 0), (k_t, 1 << 14)])], 0, True)
    batch = 64
    A = torch.empty(batch * 4096, dtype=torch.int32, device="cuda"); B = torch.empty_like(A); Cd = torch.empty_like(A)
    ops.synth_fill(A, a.raw_min, a.raw_max, 9, 0); ops.synth_fill(B, a.raw_min, a.raw_max, 9, 1)
    ops.Qgemul_poly(Cd, A, B, plan, poly, batch=batch)                                # builds the kernel
    lib = abi.load()
    lib.qb200_launch_count_reset()
    ops.Qgemul_poly(Cd, A, B, plan, poly, batch=batch)
    assert lib.qb200_launch_count() == 1


@pytest.mark.parametrize("which", ["a", "b", "ab"])
def test_gemul_with_polynomial_prologue_vs_oracle(which):
    """SURVEY 8(f) rank 3, the prologue side: Qgemul(C, Qapprox(A), Qapprox(B)) with the piecewise polynomial applied to the
    operand elements as they are staged (register tiles: in the conversion pass, once per element; thread per output: on
    load).  Same bits as the oracle's Qapprox over the operand followed by the oracle's Qgemul; device and host buffers, the
    config-3 types at 64x64x64 (one launch), a ragged shape, and the unfused runtime-descriptor path."""
    a = QType(7, 8); mul = QType(14, 8, True, RND_CONV, SAT_TCPL)
    c = QType(17, 4, True, RND_POS_INF, SAT_TCPL)
    k_t = QType(1, 14, True, RND_POS_INF, SAT_TCPL)
    segs = [(-20.0, [(k_t, 1 << 10)]), (0.0, [(k_t, 0), (k_t, 1 << 12)]), (1e9, [(k_t, 0), (k_t, 1 << 14), (k_t, -5)])]
    lib = abi.load()
    for M, N, K, batch in ((64, 64, 64, 9), (10, 6, 12, 5)):
        nlev = max(1, (K - 1).bit_length())
        lv = [QType(15 + l, 8 - min(l, 8), True, RND_POS_INF, SAT_TCPL) for l in range(nlev)]
        plan = ops.make_gemul_plan(a, a, c, M, N, K, mul=mul, add=lv)
        pa = ops.make_poly_plan(a, segs, 0, True) if "a" in which else None
        pb = ops.make_poly_plan(a, segs[1:], 0, True) if "b" in which else None
        A = torch.empty(batch * M * K, dtype=torch.int32, device="cuda"); B = torch.empty(batch * K * N, dtype=torch.int32, device="cuda")
        ops.synth_fill(A, a.raw_min, a.raw_max, 6, 0); ops.synth_fill(B, a.raw_min, a.raw_max, 6, 1)
        A[:7] = a.raw_min; B[:7] = a.raw_min; A[7:14] = a.raw_max
        An, Bn = A.cpu().numpy(), B.cpu().numpy()
        Ap, Bp = An, Bn
        if pa is not None:
            pa.len = An.size; Ap = oracle.poly_batched(pa, An)
        if pb is not None:
            pb.len = Bn.size; Bp = oracle.poly_batched(pb, Bn)
        want = oracle.gemul_batched(plan, Ap, Bp, batch)
        Cd = torch.empty(batch * M * N, dtype=torch.int32, device="cuda")
        ops.Qgemul_prologue(Cd, A, B, plan, pa, pb, batch=batch)
        assert np.array_equal(Cd.cpu().numpy(), want), (M, which)
        lib.qb200_launch_count_reset()
        ops.Qgemul_prologue(Cd, A, B, plan, pa, pb, batch=batch)
        assert lib.qb200_launch_count() == 1 and np.array_equal(Cd.cpu().numpy(), want)
        Ch = np.empty(batch * M * N, dtype=np.int32)
        ops.Qgemul_prologue(Ch, An, Bn, plan, pa, pb, batch=batch)
        assert np.array_equal(Ch, want)
        C2 = torch.empty_like(Cd)                                                     # polynomial into a scratch operand, then the runtime-descriptor Qgemul
        ops.Qgemul_prologue(C2, A, B, plan, pa, pb, batch=batch, flags=abi.QB_F_NO_SPECIALIZE)
        assert torch.equal(Cd, C2)
    # a polynomial whose result type is not the operand type is refused
    bad = ops.make_poly_plan(QType(6, 8), segs, 0, True)
    with pytest.raises(Exception):
        ops.Qgemul_prologue(Cd, A, B, plan, bad, None, batch=batch)


def test_table_lookup_matches_quantized_function():
    """ANUS::Qtable<func>(x) (readme.md:66-78): out = out_t(func(x.toDouble())) for the three predefined tables."""
    x_t = QType(4, 8, True)
    out_t = QType(6, 9, True, RND_POS_INF, SAT_TCPL)
    xs = torch.arange(x_t.raw_min, x_t.raw_max + 1, dtype=torch.int32, device="cuda")
    xd = xs.cpu().numpy().astype(np.float64) * 2.0 ** (-x_t.f)
    for func in (lambda v: 1.0 / np.sqrt(v), lambda v: 1.0 / v, np.sqrt):
        got = ops.Qtable(xs, x_t, out_t, func)
        with np.errstate(all="ignore"):
            vals = np.asarray(func(xd), dtype=np.float64)
        vals = np.where(np.isfinite(vals), vals, 0.0)
        want = oracle.quantize_f64(out_t, vals)
        assert np.array_equal(got.cpu().numpy(), want)
