"""SURVEY.md 8(f) rank 1 — the steps on either side of the path in an accuracy sweep, on the GPU:
double -> Qu (bit-exact vs the reference's Qu_s(double) goldens and the oracle), Qu -> double (exact),
and SQNR / max-abs error statistics against a float64 reference."""
import json
import os
import struct

import numpy as np
import pytest

from oracle import oracle
from qublas_b200 import ops
from qublas_b200.qtypes import QType, RND_CONV, RND_POS_INF, SAT_TCPL
from tests.helpers import load_cases, qt

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def test_quantize_matches_reference_goldens():
    n = 0
    for c in load_cases("ref_fromdouble.json"):
        t = qt(c["to"])
        x = np.array([struct.unpack("<d", struct.pack("<q", i[0]))[0] for i in c["inputs"]], dtype=np.float64)
        got = ops.quantize(torch.from_numpy(x).cuda(), t).cpu().numpy()
        assert [int(v) for v in got] == [o[0] for o in c["outputs"]], c["to"]
        n += len(x)
    assert n >= 1200


def test_quantize_64bit_wide_targets_match_reference_goldens():
    """Qu_s(double) into 64-bit-wide targets, |raw| around 2^62..2^64 and far beyond (the overflow decision must
    be made on the exact value): reference goldens outside the reference's own defect band, the oracle everywhere."""
    from tests.helpers import fd64_defect_band
    n = 0
    for c in load_cases("ref_fromdouble64.json"):
        t = qt(c["to"])
        x = np.array([struct.unpack("<d", struct.pack("<q", i[0]))[0] for i in c["inputs"]], dtype=np.float64)
        got = [int(v) for v in ops.quantize(torch.from_numpy(x).cuda(), t).cpu().numpy()]
        assert got == [int(v) for v in oracle.quantize_f64(t, x)], c["to"]
        for v, g, o in zip(x, got, c["outputs"]):
            if not fd64_defect_band(float(v), t):
                assert g == o[0], (c["to"], v)
                n += 1
    assert n >= 3200
    t = QType(63, 0)                                   # the ADVICE example: 2^62 is representable, not saturated
    assert int(ops.quantize(torch.tensor([2.0 ** 62, -2.0 ** 63], dtype=torch.float64).cuda(), t).cpu()[0]) == 2 ** 62
    assert int(ops.quantize(torch.tensor([-2.0 ** 63], dtype=torch.float64).cuda(), t).cpu()[0]) == -2 ** 63


@pytest.mark.parametrize("mem", ["device", "host"])
def test_quantize_dequantize_vs_oracle_all_modes(mem):
    rng = np.random.default_rng(12)
    for q in range(7):
        for o in range(4):
            t = QType(int(rng.integers(2, 20)), int(rng.integers(-2, 24)), bool(rng.random() < 0.8), q, o)
            if t.i + t.f < 1:
                continue
            scale = 2.0 ** t.i
            x = np.concatenate([rng.uniform(-1.5 * scale, 1.5 * scale, 4000), rng.normal(0, scale / 8, 4000),
                                (rng.integers(-4096, 4096, 500) + 0.5) * 2.0 ** (-t.f), [0.0, -0.0, 1e300, -1e300, 1e-300, np.inf, np.nan]])
            want = oracle.quantize_f64(t, x)
            xin = torch.from_numpy(x).cuda() if mem == "device" else x
            got = ops.quantize(xin, t)
            gotn = got.cpu().numpy() if mem == "device" else got
            assert np.array_equal(gotn, want), (t,)
            back = ops.dequantize(got, t)
            backn = back.cpu().numpy() if mem == "device" else back
            assert np.array_equal(backn, oracle.dequantize_f64(t, want))


def test_error_stats_of_a_quantized_gemul_vs_float64():
    """A 64x64x64 config-3 Qgemul against the float64 GEMM of the dequantized operands: SQNR and max error
    computed on the device equal the oracle's (summation order differs: 1e-12 relative tolerance)."""
    from qublas_b200.qtypes import RND_CONV, RND_POS_INF
    a = QType(7, 8); mul = QType(14, 8, True, RND_CONV, SAT_TCPL)
    lv = [QType(15 + l, 8 - l, True, RND_POS_INF, SAT_TCPL) for l in range(6)]
    c = QType(17, 4, True, RND_POS_INF, SAT_TCPL)
    plan = ops.make_gemul_plan(a, a, c, 64, 64, 64, mul=mul, add=lv)
    batch = 32
    rng = np.random.default_rng(1)
    Ad = torch.from_numpy(rng.normal(0, 20, batch * 4096)).cuda(); Bd = torch.from_numpy(rng.normal(0, 20, batch * 4096)).cuda()
    A = ops.quantize(Ad, a); B = ops.quantize(Bd, a)
    Cq = torch.empty(batch * 4096, dtype=torch.int32, device="cuda")
    ops.Qgemul(Cq, A, B, plan, batch=batch)
    Af = ops.dequantize(A, a).view(batch, 64, 64).transpose(1, 2); Bf = ops.dequantize(B, a).view(batch, 64, 64).transpose(1, 2)
    ref = torch.bmm(Af, Bf).transpose(1, 2).contiguous().view(-1)        # column-major like C
    st = ops.error_stats(Cq, ref, c)
    se, sr, mx = oracle.error_stats(c, Cq.cpu().numpy(), ref.cpu().numpy())
    assert abs(st["sum_sq_err"] - se) <= 1e-12 * se and abs(st["sum_sq_ref"] - sr) <= 1e-12 * sr and st["max_abs_err"] == mx
    assert 40.0 < st["sqnr_db"] < 120.0 and st["n"] == batch * 4096
    again = ops.error_stats(Cq, ref, c)
    assert again["sum_sq_err"] == st["sum_sq_err"]                         # fixed reduction order: reproducible


@pytest.mark.parametrize("M,N,K,ta,tb,wb", [(64, 64, 64, False, False, 0), (32, 48, 40, True, False, 0), (16, 8, 24, False, True, 2), (10, 6, 12, False, False, 0),
                                             (64, 64, 64, True, True, 2)])
def test_gemul_error_in_one_pass_equals_reference_gemm_plus_error_stats(M, N, K, ta, tb, wb):
    """qb200_gemul_error_batched (the accuracy sweep in one pass: the float64 reference of each instance lives in registers)
    against the two-step path -- qb200_gemul_f64_batched + qb200_error_stats -- and against numpy: the same reference values
    (same k-ascending FMA chain), sums equal up to the summation order, the maximum error exactly; host and device memory;
    int16 words; a shape that is not a multiple of 4 (falls back to the two kernels inside the same call)."""
    from qublas_b200.qtypes import RND_CONV, RND_POS_INF
    a = QType(7, 8); mul = QType(14, 8, True, RND_CONV, SAT_TCPL)
    nl = max(1, (K - 1).bit_length())
    lv = [QType(15 + l, 8, True, RND_POS_INF, SAT_TCPL) for l in range(nl)]
    c = QType(17, 4, True, RND_POS_INF, SAT_TCPL)
    plan = ops.make_gemul_plan(a, a, c, M, N, K, mul=mul, add=lv, trans_a=ta, trans_b=tb, word_bytes=(wb, wb, 0))
    batch = 37
    rng = np.random.default_rng(M * N + K)
    dt = np.int16 if wb == 2 else np.int32
    A = rng.integers(a.raw_min, a.raw_max + 1, size=batch * M * K).astype(dt); B = rng.integers(a.raw_min, a.raw_max + 1, size=batch * K * N).astype(dt)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    Cq = torch.empty(batch * M * N, dtype=torch.int32, device="cuda")
    ops.Qgemul(Cq, dA, dB, plan, batch=batch)
    one = ops.gemul_error(Cq, dA, dB, plan, batch=batch)
    ref = ops.gemul_f64(dA, dB, plan, batch=batch)
    two = ops.error_stats(Cq, ref, c)
    assert one["n"] == batch * M * N and one["max_abs_err"] == two["max_abs_err"]
    assert abs(one["sum_sq_err"] - two["sum_sq_err"]) <= 1e-12 * two["sum_sq_err"] and abs(one["sum_sq_ref"] - two["sum_sq_ref"]) <= 1e-12 * two["sum_sq_ref"]
    # numpy yardstick (exact: every partial sum is far below 2^53)
    Am = A.astype(np.float64).reshape(batch, M, K) if ta else A.astype(np.float64).reshape(batch, K, M).transpose(0, 2, 1)
    Bm = B.astype(np.float64).reshape(batch, K, N) if tb else B.astype(np.float64).reshape(batch, N, K).transpose(0, 2, 1)
    want = (Am @ Bm).transpose(0, 2, 1).reshape(-1) * 2.0 ** (-2 * a.f)
    err = Cq.cpu().numpy().astype(np.float64) * 2.0 ** (-c.f) - want
    assert one["max_abs_err"] == np.abs(err).max() and abs(one["sum_sq_err"] - (err * err).sum()) <= 1e-11 * (err * err).sum()
    # reproducible, and the same from host memory
    assert ops.gemul_error(Cq, dA, dB, plan, batch=batch) == one
    assert ops.gemul_error(Cq.cpu().numpy(), A, B, plan, batch=batch) == one
