"""The algebra the config-3 register-tile kernel relies on (qublas_b200/csrc/qb_fast_gemul.cuh, qb_device.cuh), restated in
numpy and checked against the CPU oracle -- no GPU involved.  What the kernel does differently from the reference's
integer shift/mask/clamp formulation is (1) Qmul<RND::CONV> as ONE fp32 FMA with a magic constant, (2) the saturation of
min x min folded into substituted operands, (3) the tree carried as pre-shift sums, negated, with the FMA bias removed once
at the root.  Each identity is exercised on exactly the operand sets where it could break (substituted values against every
partner, ties, extremes) and on a whole 64x64x64 instance; the GPU tests sweep all 2^32 pairs through the real device code."""
import numpy as np
import pytest

from oracle import oracle
from qublas_b200 import ops
from qublas_b200.qtypes import QType, RND_CONV, RND_POS_INF, SAT_TCPL


def fma_rn_f32(a, b, c):
    """fp32 fused multiply-add, round to nearest even: the product and the sum are exact in float64 for these magnitudes
    (|a*b| < 2^31, c a multiple of 2^8 below 2^33), so one rounding float64 -> float32 is the FMA's single rounding."""
    return (a.astype(np.float64) * b.astype(np.float64) + np.float64(c)).astype(np.float32)


def qmul_oracle(a_t, b_t, mul_t, x, y):
    """Qmul<mul_t>(x, y) through the oracle: a batch of 1x1x1 Qgemul instances whose tree and C conversion are identities."""
    wide = QType(mul_t.i + 2, mul_t.f, True, RND_POS_INF, SAT_TCPL)
    plan = ops.make_gemul_plan(a_t, b_t, wide, 1, 1, 1, mul=mul_t, add=[wide])
    return oracle.gemul_batched(plan, x.astype(np.int32), y.astype(np.int32), x.size).astype(np.int64)


CASES = [  # operand types, product type (RND::CONV, SAT::TCPL): hi = 2^(pa + pb - S) - 1, i.e. only min x min saturates
    (QType(7, 8), QType(7, 8), QType(14, 8, True, RND_CONV, SAT_TCPL)),       # config 3: 16 x 16 bits, S = 8
    (QType(3, 4), QType(5, 6), QType(8, 6, True, RND_CONV, SAT_TCPL)),        # 8 x 12 bits, S = 4
    (QType(7, 8), QType(7, 8), QType(14, 4, True, RND_CONV, SAT_TCPL)),       # 16 x 16 bits, S = 12
    (QType(9, 2), QType(2, 9), QType(11, 0, True, RND_CONV, SAT_TCPL)),       # 12 x 12 bits, S = 11 = log2(-min)
]


@pytest.mark.parametrize("a_t,b_t,mul_t", CASES)
def test_fma_product_and_operand_substitution_equal_the_integer_quantizer(a_t, b_t, mul_t):
    S = a_t.f + b_t.f - mul_t.f
    pa, pb = (-a_t.raw_min).bit_length() - 1, (-b_t.raw_min).bit_length() - 1
    assert pa >= S and pb >= S and mul_t.raw_max == (1 << (pa + pb - S)) - 1 and pa + pb - S + 1 <= 24      # fp_subst's conditions
    Cm = 3 * 2.0 ** (22 + S)                                                  # 1.5 * 2^(23+S)
    kbits = np.int64(((150 + S) << 23) | 0x400000)
    a_sub = np.float32(-(2.0 ** pa - 2.0 ** (S - pb - 1))); b_sub = np.float32(-(2.0 ** pb - 2.0 ** (S - pa - 1)))
    assert float(a_sub) == -(2.0 ** pa - 2.0 ** (S - pb - 1))                 # exactly representable
    rng = np.random.default_rng(S)
    xs = np.arange(a_t.raw_min, a_t.raw_max + 1, dtype=np.int64); ys = np.arange(b_t.raw_min, b_t.raw_max + 1, dtype=np.int64)
    # the pairs where the substitution acts: min(a) with every b, every a with min(b); plus ties, extremes and a random sample
    px = [np.full(ys.size, a_t.raw_min), xs, rng.integers(a_t.raw_min, a_t.raw_max + 1, 200000), np.array([a_t.raw_max, a_t.raw_max, a_t.raw_min + 1])]
    py = [ys, np.full(xs.size, b_t.raw_min), rng.integers(b_t.raw_min, b_t.raw_max + 1, 200000), np.array([b_t.raw_max, b_t.raw_min, b_t.raw_min + 1])]
    x = np.concatenate(px); y = np.concatenate(py)
    tie = (x * y) % (1 << S) == (1 << (S - 1))
    assert tie.sum() > 100                                                     # RND::CONV ties are in the set
    want = qmul_oracle(a_t, b_t, mul_t, x, y)
    # (1) plain FMA form + integer clamp on the biased bits
    bits = fma_rn_f32(x.astype(np.float32), y.astype(np.float32), Cm).view(np.int32).astype(np.int64)
    assert np.array_equal(np.minimum(bits - kbits, mul_t.raw_max), want)
    # (2) substituted operands, NO clamp
    xf = np.maximum(x.astype(np.float32), a_sub); yf = np.maximum(y.astype(np.float32), b_sub)
    bits2 = fma_rn_f32(xf, yf, Cm).view(np.int32).astype(np.int64)
    assert np.array_equal(bits2 - kbits, want)
    assert want.max() == mul_t.raw_max                                          # the saturated value was produced
    # (3) negated A (the negated S form): odd symmetry of round-to-nearest-even
    bits3 = fma_rn_f32(-xf, yf, Cm).view(np.int32).astype(np.int64)
    assert np.array_equal(-(bits3 - kbits), want)


def test_negated_deferred_shift_tree_equals_the_reference_tree_on_a_whole_instance():
    """Config 3, one 64x64x64 instance full of extremes: products as biased float bit patterns of (-a') * b', groups of four
    summed WITHOUT removing the bias, merges S' = (Sa >> 1) + (Sb >> 1) with no rounding constant, D / 2 - (S_root >> 1) at the
    root, C's converting assignment -- against the oracle's Qmul -> per-level Qadd tree -> conversion."""
    a = QType(7, 8); mul = QType(14, 8, True, RND_CONV, SAT_TCPL)
    lv = [QType(15 + l, 8 - l, True, RND_POS_INF, SAT_TCPL) for l in range(6)]
    c = QType(17, 4, True, RND_POS_INF, SAT_TCPL)
    M = N = K = 64
    plan = ops.make_gemul_plan(a, a, c, M, N, K, mul=mul, add=lv)
    rng = np.random.default_rng(11)
    A = rng.integers(a.raw_min, a.raw_max + 1, M * K).astype(np.int32); B = rng.integers(a.raw_min, a.raw_max + 1, K * N).astype(np.int32)
    A[rng.random(A.size) < 0.1] = a.raw_min; B[rng.random(B.size) < 0.1] = a.raw_min; A[rng.random(A.size) < 0.05] = a.raw_max
    want = oracle.gemul_batched(plan, A, B, 1).astype(np.int64).reshape(N, M)          # column-major: [j][i]
    S = 8; Cm = 3 * 2.0 ** (22 + S); kbits = ((150 + S) << 23) | 0x400000
    sub = np.float32(-(2.0 ** 15 - 2.0 ** -8))
    Am = -np.maximum(A.astype(np.float32), sub).reshape(K, M)                            # [k][i], negated
    Bm = np.maximum(B.astype(np.float32), sub).reshape(N, K)                             # [j][k]
    # leaves: biased bit patterns as uint32, [j][i][k]
    bits = fma_rn_f32(Am.T[None, :, :], Bm[:, None, :], Cm).view(np.int32).astype(np.int64) & 0xffffffff
    g = bits.reshape(N, M, K // 4, 4).sum(axis=3) & 0xffffffff                           # level 0 exact + level 1 pre-shift sum, bias 4 * kbits kept
    Sv = np.where(g >= 1 << 31, g - (1 << 32), g)                                        # as int32
    D = (4 * kbits) & 0xffffffff
    assert D < 1 << 31 and D % 2 == 0
    while Sv.shape[2] > 1:                                                               # merges: arithmetic shifts, no constant
        Sv = (Sv[:, :, 0::2] >> 1) + (Sv[:, :, 1::2] >> 1)
        assert np.abs(Sv).max() < 1 << 31
    root = D // 2 - (Sv[:, :, 0] >> 1)                                                   # un-negate, remove the bias once
    assert lv[-1].f == 3 and c.f == 4
    out = np.clip(root << 1, c.raw_min, c.raw_max)                                       # level 5 has 3 fraction bits, C has 4: exact left shift, then SAT::TCPL
    assert np.array_equal(out, want)
