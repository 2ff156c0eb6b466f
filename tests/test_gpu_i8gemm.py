"""BASELINE config 5: Qgemul with int8-range operands and lossless wide accumulators on the
tcgen05 kind::i8 tensor-core path.  Bit-exact against the oracle (small shapes), against the
integer-ALU kernel (same ABI call with QB_F_NO_TENSOR_CORE) and against an exact float64 GEMM."""
import numpy as np
import pytest

from oracle import oracle
from qublas_b200 import abi, ops
from qublas_b200.qtypes import QType, RND_POS_INF, SAT_TCPL

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def c5_plan(M, N, K, c=None, trans_a=False, trans_b=False, full_prec=False):
    a = QType(7, 0)
    acc = QType(30, 0)
    mul = {"full": True} if full_prec else QType(15, 0)
    return a, ops.make_gemul_plan(a, a, c or acc, M, N, K, mul=mul, add=[acc], trans_a=trans_a, trans_b=trans_b)


@pytest.mark.parametrize("M,N,K,ta,tb", [(128, 256, 128, False, False), (256, 512, 256, False, False), (128, 256, 384, True, False),
                                         (256, 256, 128, False, True), (384, 256, 256, True, True)])
def test_i8_tensor_core_vs_oracle(M, N, K, ta, tb):
    rng = np.random.default_rng(M + N + K)
    a, plan = c5_plan(M, N, K, trans_a=ta, trans_b=tb)
    info = ops.plan_info_gemul(plan)
    assert info.tensor_core == 1 and info.lossless == 1 and b"tcgen05" in info.kernel
    A = rng.integers(-128, 128, size=M * K).astype(np.int32)
    B = rng.integers(-128, 128, size=K * N).astype(np.int32)
    A[:7] = -128; B[:7] = -128                                  # min x min products
    want = oracle.gemul_batched(plan, A, B, 1)
    Cd = torch.empty(M * N, dtype=torch.int32, device="cuda")
    ops.Qgemul(Cd, torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), plan)
    assert np.array_equal(Cd.cpu().numpy(), want)
    C2 = torch.empty_like(Cd)
    ops.Qgemul(C2, torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), plan, flags=abi.QB_F_NO_TENSOR_CORE)
    assert torch.equal(Cd, C2)


def test_i8_epilogue_quantizes_once_into_c_type():
    """C narrower than the accumulator: the single quantizer left is the converting assignment."""
    rng = np.random.default_rng(3)
    M, N, K = 128, 256, 256
    c = QType(12, -2, True, RND_POS_INF, SAT_TCPL)              # drops 2 integer LSBs and saturates
    a, plan = c5_plan(M, N, K, c=c)
    assert ops.plan_info_gemul(plan).tensor_core == 1
    A = rng.integers(-128, 128, size=M * K).astype(np.int32)
    B = rng.integers(-128, 128, size=K * N).astype(np.int32)
    want = oracle.gemul_batched(plan, A, B, 1)
    Cd = torch.empty(M * N, dtype=torch.int32, device="cuda")
    ops.Qgemul(Cd, torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), plan)
    assert np.array_equal(Cd.cpu().numpy(), want)
    assert (np.abs(want) == c.raw_max).any() or (want == c.raw_min).any()      # saturation was exercised


def test_fullprec_product_is_not_lossless_and_stays_on_alu():
    """Qmul<FullPrec> of two 7.0 types is 14.0 and cannot hold (-128)*(-128): the predicate
    must refuse the tensor-core path and the ALU path must reproduce the saturated product."""
    M, N, K = 128, 256, 128
    a, plan = c5_plan(M, N, K, full_prec=True)
    info = ops.plan_info_gemul(plan)
    assert info.tensor_core == 0 and info.lossless == 0
    A = np.full(M * K, -128, dtype=np.int32); B = np.full(K * N, -128, dtype=np.int32)
    want = oracle.gemul_batched(plan, A, B, 1)
    Cd = torch.empty(M * N, dtype=torch.int32, device="cuda")
    ops.Qgemul(Cd, torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), plan)
    assert np.array_equal(Cd.cpu().numpy(), want) and int(want[0]) == K * 16383


def test_i8_large_vs_exact_float64_gemm():
    """2048^3 (full tiles, many k-blocks, several waves): all partial sums < 2^53, so a float64 GEMM is exact."""
    M = N = K = 2048
    a, plan = c5_plan(M, N, K)
    A = torch.empty(M * K, dtype=torch.int32, device="cuda"); B = torch.empty(K * N, dtype=torch.int32, device="cuda")
    ops.synth_fill(A, -128, 127, 5, 0); ops.synth_fill(B, -128, 127, 5, 1)
    Cd = torch.empty(M * N, dtype=torch.int32, device="cuda")
    ops.Qgemul(Cd, A, B, plan)
    Am = A.view(K, M).t().double(); Bm = B.view(N, K).t().double()         # column-major views
    ref = (Am @ Bm).t().contiguous().view(-1)                               # back to column-major
    assert torch.equal(Cd.double(), ref)


@pytest.mark.parametrize("M,N,K,ta,tb", [(300, 500, 200, False, False), (380, 500, 333, True, False), (600, 500, 250, False, True), (250, 700, 380, True, True)])
def test_i8_tensor_core_ragged_shapes_run_on_zero_padded_operands(M, N, K, ta, tb):
    """Shapes that are not tile multiples: the operands are packed into zero-padded int8 staging buffers (zeros add
    nothing to an exact dot product) and the epilogue stores only the real M x N block."""
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    a, plan = c5_plan(M, N, K, trans_a=ta, trans_b=tb)
    info = ops.plan_info_gemul(plan)
    assert info.tensor_core == 1 and b"tcgen05" in info.kernel
    A = rng.integers(-128, 128, size=M * K).astype(np.int32)
    B = rng.integers(-128, 128, size=K * N).astype(np.int32)
    want = oracle.gemul_batched(plan, A, B, 1)
    guard = 64
    Cd = torch.full((M * N + guard,), 123456789, dtype=torch.int32, device="cuda")        # canary behind the result
    ops.Qgemul(Cd[:M * N], torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), plan)
    assert np.array_equal(Cd[:M * N].cpu().numpy(), want)
    assert bool((Cd[M * N:] == 123456789).all())
    # small or badly padded ragged shapes stay on the ALU kernels
    assert ops.plan_info_gemul(c5_plan(40, 50, 60)[1]).tensor_core == 0
    assert ops.plan_info_gemul(c5_plan(129, 4096, 4096)[1]).tensor_core == 1           # 256/129 < 2
    assert ops.plan_info_gemul(c5_plan(5, 4096, 4096)[1]).tensor_core == 0


@pytest.mark.parametrize("M,N,K", [(256, 256, 256), (384, 512, 128), (500, 500, 240), (8192, 512, 1024)])
def test_i8_operands_handed_over_k_major_are_read_in_place(M, N, K):
    """int8 words, K contiguous (A stored K x M = trans_a, B stored K x N) -> the tensor maps sit on the caller's
    buffers, no pack kernel runs (ONE launch per call); ragged extents rely on TMA's zero fill.  Same bits as the oracle and
    as the int32-word call; a buffer that is not 16-byte aligned takes the pack path and still gives the same bits."""
    rng = np.random.default_rng(M * 3 + N + K)
    a = QType(7, 0); acc = QType(30, 0)
    plan8 = ops.make_gemul_plan(a, a, acc, M, N, K, mul=QType(15, 0), add=[acc], trans_a=True, word_bytes=(1, 1, 0))
    plan32 = ops.make_gemul_plan(a, a, acc, M, N, K, mul=QType(15, 0), add=[acc], trans_a=True)
    assert ops.plan_info_gemul(plan8).tensor_core == 1
    A = rng.integers(-128, 128, size=M * K).astype(np.int8); B = rng.integers(-128, 128, size=K * N).astype(np.int8)
    A[:5] = -128; B[:5] = -128
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    Cd = torch.empty(M * N, dtype=torch.int32, device="cuda")
    lib = abi.load()
    ops.Qgemul(Cd, dA, dB, plan8)                                # first call: function attributes etc.
    torch.cuda.synchronize(); lib.qb200_launch_count_reset()
    ops.Qgemul(Cd, dA, dB, plan8)
    torch.cuda.synchronize()
    assert lib.qb200_launch_count() == 1                         # the GEMM alone: no pack kernels
    C32 = torch.empty_like(Cd)
    lib.qb200_launch_count_reset()
    ops.Qgemul(C32, dA.to(torch.int32), dB.to(torch.int32), plan32)
    torch.cuda.synchronize()
    fused = M % 256 == 0 and N % 256 == 0 and K % 128 == 0       # int32 words, K contiguous: narrowed inside the GEMM kernel
    assert lib.qb200_launch_count() == (1 if fused else 3) and torch.equal(Cd, C32)
    if M * N * K <= 1 << 26:
        assert np.array_equal(Cd.cpu().numpy(), oracle.gemul_batched(plan8, A, B, 1))
    else:
        rows = torch.arange(0, M, 61, device="cuda")
        ref = dA.view(M, K)[rows].double() @ dB.view(N, K).t().double()
        assert torch.equal(Cd.view(N, M).t()[rows].double(), ref)
    # word-aligned only: staged through the pack kernels
    bufA = torch.zeros(M * K + 16, dtype=torch.int8, device="cuda"); bufA[1:1 + M * K] = dA
    C3 = torch.empty_like(Cd)
    ops.Qgemul(C3, bufA[1:1 + M * K], dB, plan8)
    assert torch.equal(C3, Cd)


@pytest.mark.parametrize("M,N,K,ta", [(256, 256, 128, True), (512, 768, 2176, True), (1024, 512, 1024, False), (4096, 2048, 3072, True), (256, 512, 2176, False),
                                      (2048, 2048, 2048, False)])
def test_i8_int32_words_are_narrowed_inside_the_gemm_kernel(M, N, K, ta):
    """VERDICT r1 task 5: the reference's int32 words with K contiguous are narrowed to int8 by packer warps of the GEMM
    kernel while the tensor cores run (k_i8gemm_pair_fused): ONE launch, whether A is stored K x M (straight narrowing) or
    M x K (the reference's natural layout: transposed in registers by the same warps).  K = 2176 ends in a partial k-chunk;
    the call is repeated (the flags are cleared per call) and also run on a batch of two instances."""
    rng = np.random.default_rng(M + 5 * N + K)
    a = QType(7, 0); acc = QType(30, 0)
    plan = ops.make_gemul_plan(a, a, acc, M, N, K, mul=QType(15, 0), add=[acc], trans_a=ta)
    assert ops.plan_info_gemul(plan).tensor_core == 1
    nb = 2
    A = rng.integers(-128, 128, size=nb * M * K).astype(np.int32); B = rng.integers(-128, 128, size=nb * K * N).astype(np.int32)
    A[:9] = -128; B[:9] = -128
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    Cd = torch.empty(nb * M * N, dtype=torch.int32, device="cuda")
    lib = abi.load()
    ops.Qgemul(Cd[:M * N], dA[:M * K], dB[:K * N], plan)
    torch.cuda.synchronize(); lib.qb200_launch_count_reset()
    for _ in range(3):
        Cd.zero_()
        ops.Qgemul(Cd, dA, dB, plan, batch=nb)
    torch.cuda.synchronize()
    assert lib.qb200_launch_count() == 3 * nb                   # one launch per instance, whichever way A is stored
    for i in range(nb):
        Ai, Bi = dA[i * M * K:(i + 1) * M * K], dB[i * K * N:(i + 1) * K * N]
        Am = (Ai.view(M, K) if ta else Ai.view(K, M).t()).double()
        ref = (Am @ Bi.view(N, K).t().double()).t().contiguous().view(-1)           # exact: every partial sum < 2^53
        assert torch.equal(Cd[i * M * N:(i + 1) * M * N].double(), ref)
    if M * N * K <= 1 << 25:
        assert np.array_equal(Cd[:M * N].cpu().numpy(), oracle.gemul_batched(plan, A[:M * K], B[:K * N], 1))
