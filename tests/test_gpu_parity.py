"""GPU parity tests: the CUDA path (through the C ABI of libqb200.so) against the CPU oracle and the
reference-generated golden fixtures.  Integer work => bit-exact, no tolerance anywhere."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle
from qublas_b200 import abi, ops
from qublas_b200.qtypes import (QType, RND_CONV, RND_POS_INF, SAT_TCPL, TREE_VARIADIC, TREE_VECTOR, mul_merge,
                                resolve_tf, tree_out_type)
from tests.helpers import (cgemul_plan, form_id, gemul_plan, load_cases, np_words, qt, reduce_plan, tf_tags,
                           tfcmul_plan)

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if abi.load().qb200_device_count() < 1:
        pytest.fail("no CUDA device: the GPU tests must run on the B200 box")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rand_words(rng, t: QType, n):
    return rng.integers(t.raw_min, t.raw_max + 1, size=n, dtype=np.int64).astype(oracle.np_dtype(t.word_bytes))


# ------------------------------------------------------------------------------------------------
# golden fixtures produced by the unmodified reference header
# ------------------------------------------------------------------------------------------------
def test_golden_converting_ctor_and_qmul():
    """Scalar goldens on the GPU: conv cases through qb200_convert, mul cases through a 1x1x1 Qgemul."""
    n = 0
    for c in load_cases("ref_scalar.json"):
        a, b = qt(c["a"]), qt(c["b"])
        if c["op"] == "conv":
            to = qt(c["to"])
            x = np_words([i[0] for i in c["inputs"]], a)
            got = ops.convert(dev(x), a, to).cpu().numpy()
            assert [int(v) for v in got] == [o[0] for o in c["outputs"]], c
            n += len(x)
        elif c["op"] == "mul":
            to = qt(c["to"]) if "to" in c else mul_merge(a, b, c.get("tags"))
            plan = ops.make_gemul_plan(a, b, to, 1, 1, 1, mul=to)
            A = np_words([i[0] for i in c["inputs"]], a); B = np_words([i[1] for i in c["inputs"]], b)
            Cd = torch.empty(len(A), dtype=torch.int32 if to.word_bytes == 4 else torch.int64, device="cuda")
            ops.Qgemul(Cd, dev(A), dev(B), plan, batch=len(A))
            assert [int(v) for v in Cd.cpu().numpy()] == [o[0] for o in c["outputs"]], c
            n += len(A)
    assert n > 5000


@pytest.mark.parametrize("mem", ["device", "host"])
def test_golden_reduce(mem):
    for c in load_cases("ref_reduce.json"):
        elem = qt(c["elem"]); lv = [qt(t) for t in c["levels"]]
        x = np_words(sum(c["inputs"], []), elem)
        batch = len(c["inputs"])
        xin = dev(x) if mem == "device" else x
        out, ot = ops.Qreduce(xin, elem, lv, length=c["len"], batch=batch, form=form_id(c["form"]))
        got = out.cpu().numpy() if mem == "device" else out
        assert [int(v) for v in got] == [o[0] for o in c["outputs"]], (c["form"], c["len"], c["elem"], c["levels"])
        assert [ot.i, ot.f, int(ot.s), ot.q, ot.o] == c["outputs"][0][1:6]


def test_golden_tfcmul():
    for c in load_cases("ref_tf.json"):
        xr, xi, yr, yi = (qt(c[k]) for k in ("x_re", "x_im", "y_re", "y_im"))
        wx = max(xr.word_bytes, xi.word_bytes); wy = max(yr.word_bytes, yi.word_bytes)
        x = np.array([i[0:2] for i in c["inputs"]], dtype=oracle.np_dtype(wx)).reshape(-1)
        y = np.array([i[2:4] for i in c["inputs"]], dtype=oracle.np_dtype(wy)).reshape(-1)
        out, _ = ops.tfcmul(dev(x), dev(y), xr, xi, yr, yi, tf_tags(c["tags"]))
        got = out.cpu().numpy().reshape(-1, 2)
        assert [[int(a), int(b)] for a, b in got] == [o[0:2] for o in c["outputs"]], c["tags"]


@pytest.mark.parametrize("mem", ["device", "host"])
def test_golden_gemul_readme_and_random(mem):
    for c in load_cases("ref_gemul.json"):
        a, b, mul, cc = (qt(c[k]) for k in ("a", "b", "mul", "c"))
        lv = [qt(t) for t in c["levels"]]
        M, N, K = c["M"], c["N"], c["K"]
        plan = gemul_plan(a, b, mul, cc, lv, M, N, K, trans_a=c["trans_a"], form=form_id(c["form"]))
        batch = len(c["inputs"])
        A = np_words(sum([i[:M * K] for i in c["inputs"]], []), a)
        B = np_words(sum([i[M * K:] for i in c["inputs"]], []), b)
        if mem == "device":
            Cd = torch.empty(batch * M * N, dtype=torch.int32 if cc.word_bytes == 4 else torch.int64, device="cuda")
            ops.Qgemul(Cd, dev(A), dev(B), plan, batch=batch)
            got = Cd.cpu().numpy()
        else:
            got = np.empty(batch * M * N, dtype=oracle.np_dtype(cc.word_bytes))
            ops.Qgemul(got, A, B, plan, batch=batch)
        assert [int(v) for v in got] == sum(c["outputs"], []), (M, N, K, c["form"])


def test_golden_cgemul():
    for c in load_cases("ref_cgemul.json"):
        xr, xi, yr, yi = (qt(c[k]) for k in ("x_re", "x_im", "y_re", "y_im"))
        tags = tf_tags(c["tags"]); lv = [qt(t) for t in c["levels"]]
        K = c["K"]
        r = resolve_tf(xr, xi, yr, yi, tags)
        o_re = tree_out_type(K, r["AB"], lv, TREE_VARIADIC); o_im = tree_out_type(K, r["BC"], lv, TREE_VARIADIC)
        plan = cgemul_plan(xr, xi, yr, yi, tags, lv, lv, o_re, o_im, 1, 1, K, form=TREE_VARIADIC)
        batch = len(c["inputs"])
        v = np.array(c["inputs"], dtype=np.int32).reshape(batch, K, 4)
        A = np.ascontiguousarray(v[:, :, 0:2]).reshape(-1); B = np.ascontiguousarray(v[:, :, 2:4]).reshape(-1)
        Cd = torch.empty(batch * 2, dtype=torch.int32, device="cuda")
        ops.Qcgemul(Cd, dev(A), dev(B), plan, batch=batch)
        got = Cd.cpu().numpy().reshape(-1, 2)
        assert [[int(a), int(b)] for a, b in got] == [o[0:2] for o in c["outputs"]]


# ------------------------------------------------------------------------------------------------
# randomized plans against the oracle
# ------------------------------------------------------------------------------------------------
def _rand_levels(rng, elem: QType, n, wide=False):
    lv = []
    i, f = elem.i, elem.f
    for _ in range(n):
        i += int(rng.integers(0, 3)); f -= int(rng.integers(0, 3))
        if i + f < 1:
            f = 1 - i
        lv.append(QType(i, f, bool(rng.random() < 0.9), int(rng.integers(0, 7)), int(rng.integers(0, 4))))
    return lv


@pytest.mark.parametrize("seed", range(6))
def test_random_reduce_vs_oracle(seed):
    rng = np.random.default_rng(1000 + seed)
    for n in [2, 63, 64, 65, 255, 256, 257, 1000, 4095, 4096, 4097, 5000, 65536 + 1, 70001, 1 << 17]:
        wide = rng.random() < 0.3
        elem = QType(int(rng.integers(3, 10)) + (20 if wide else 0), int(rng.integers(2, 9)), True, int(rng.integers(0, 7)),
                     int(rng.integers(0, 4)))
        lv = _rand_levels(rng, elem, int(rng.choice([0, 1, 3, 8, 20])))
        batch = int(rng.choice([1, 3])) if n > 4096 else int(rng.choice([1, 5, 17]))
        x = rand_words(rng, elem, n * batch)
        plan = reduce_plan(elem, lv, n)
        want, _ = oracle.reduce_batched(plan, x, batch)
        try:
            out, _ = ops.Qreduce(dev(x), elem, lv, length=n, batch=batch)
        except abi.QbError as e:
            assert e.code == abi.QB_EDOMAIN        # a random plan may land on a fenced reference defect
            continue
        assert np.array_equal(out.cpu().numpy(), want), (n, elem, lv)
    # variadic form goes through the per-vector streaming kernel
    for n in [3, 5, 6, 7, 11, 100, 1001]:
        elem = QType(6, 6, True, int(rng.integers(0, 7)), int(rng.integers(0, 4)))
        lv = _rand_levels(rng, elem, 3)
        x = rand_words(rng, elem, n * 9)
        plan = reduce_plan(elem, lv, n, TREE_VARIADIC)
        want, _ = oracle.reduce_batched(plan, x, 9)
        out, _ = ops.Qreduce(dev(x), elem, lv, length=n, batch=9, form=TREE_VARIADIC)
        assert np.array_equal(out.cpu().numpy(), want), (n, elem, lv)


@pytest.mark.parametrize("n,level", [(1 << 16, 9), (70001, 8), (5000, 12), (300, 8), (256, 8), (1 << 14, 0)])
def test_reduce_partials_depthwise(n, level):
    """low-level errors are masked in the final scalar, so intermediate levels are compared too."""
    rng = np.random.default_rng(n)
    elem = QType(8, 8)
    lv = [QType(9 + l, 8 - l, True, l % 7, SAT_TCPL) for l in range(8)]
    x = rand_words(rng, elem, n * 2)
    plan = reduce_plan(elem, lv, n)
    want, wparts = oracle.reduce_batched(plan, x, 2, partial_level=level)
    out, _, parts = ops.Qreduce(dev(x), elem, lv, length=n, batch=2, partial_level=level)
    assert np.array_equal(parts.cpu().numpy(), wparts)
    assert np.array_equal(out.cpu().numpy(), want)


@pytest.mark.parametrize("seed", range(4))
def test_random_gemul_vs_oracle(seed):
    rng = np.random.default_rng(2000 + seed)
    for (M, N, K) in [(1, 1, 1), (3, 5, 7), (8, 8, 8), (16, 4, 33), (5, 5, 64), (64, 64, 64), (33, 17, 100)]:
        big = rng.random() < 0.35          # forces the 64-bit arithmetic class
        a = QType(int(rng.integers(2, 8)) + (14 if big else 0), int(rng.integers(2, 9)), True, 5, 0)
        b = QType(int(rng.integers(2, 8)), int(rng.integers(2, 9)) + (8 if big else 0), True, 5, 0)
        mul = QType(a.i + b.i - int(rng.integers(0, 3)), max(a.f, b.f) - int(rng.integers(0, 3)), True,
                    int(rng.integers(0, 7)), int(rng.integers(0, 4)))
        lv = _rand_levels(rng, mul, int(rng.choice([0, 1, 2, 6])))
        last = lv[-1] if lv else mul
        cc = QType(last.i + int(rng.integers(-2, 3)), max(last.f - int(rng.integers(0, 4)), 1 - last.i + 2), True,
                   int(rng.integers(0, 7)), int(rng.integers(0, 4)))
        form = int(rng.integers(0, 2))
        ta, tb = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        plan = gemul_plan(a, b, mul, cc, lv, M, N, K, trans_a=ta, trans_b=tb, form=form)
        batch = 3
        A = rand_words(rng, a, batch * M * K); B = rand_words(rng, b, batch * K * N)
        want = oracle.gemul_batched(plan, A, B, batch)
        Cd = torch.empty(batch * M * N, dtype=torch.int32 if cc.word_bytes == 4 else torch.int64, device="cuda")
        try:
            ops.Qgemul(Cd, dev(A), dev(B), plan, batch=batch)
        except abi.QbError as e:
            assert e.code == abi.QB_EDOMAIN
            continue
        assert np.array_equal(Cd.cpu().numpy(), want), (M, N, K, a, b, mul, lv, cc, form, ta, tb)


def _c4_plan(M, N, K):
    t = QType(3, 12); P = QType(6, 12); S = QType(7, 12)
    tags = {"abT": QType(4, 12), "cdT": QType(4, 11), "abcT": P, "cdbT": P, "badT": P, "ABT": S, "BCT": S}
    lv = [QType(8 + l, 12 - l, True, RND_CONV, SAT_TCPL) for l in range(8)]
    out_t = tree_out_type(K, S, lv, TREE_VARIADIC)
    return t, cgemul_plan(t, t, t, t, tags, lv, lv, out_t, out_t, M, N, K, form=TREE_VARIADIC)


@pytest.mark.parametrize("half_range", [True, False])
def test_cgemul_c4_matvec_vs_oracle(half_range):
    """BASELINE config 4: 256x256 complex mat-vec, TFComplexMul 3-mult/5-add, 8-level CONV tree.
    Full-range inputs saturate the un-taggable b-a node in ~1/4 of the products."""
    rng = np.random.default_rng(7)
    t, plan = _c4_plan(256, 1, 256)
    batch = 4
    lo, hi = (-(1 << 14), (1 << 14) - 1) if half_range else (t.raw_min, t.raw_max)
    H = rng.integers(lo, hi + 1, size=batch * 256 * 256 * 2).astype(np.int32)
    x = rng.integers(lo, hi + 1, size=batch * 256 * 2).astype(np.int32)
    want = oracle.cgemul_batched(plan, H, x, batch)
    y = torch.empty(batch * 256 * 2, dtype=torch.int32, device="cuda")
    ops.Qcgemul(y, dev(H), dev(x), plan, batch=batch)
    assert np.array_equal(y.cpu().numpy(), want)


def test_random_cgemul_vs_oracle():
    rng = np.random.default_rng(11)
    for (M, N, K, form) in [(2, 3, 5, TREE_VARIADIC), (4, 4, 16, TREE_VECTOR), (3, 1, 7, TREE_VECTOR), (5, 2, 12, TREE_VARIADIC)]:
        def rt():
            return QType(int(rng.integers(2, 7)), int(rng.integers(1, 8)), True, int(rng.integers(0, 7)), int(rng.integers(0, 4)))
        ar, ai, br, bi = rt(), rt(), rt(), rt()
        tags = {k: QType(int(rng.integers(5, 12)), int(rng.integers(2, 10)), True, int(rng.integers(0, 7)), int(rng.integers(0, 4)))
                for k in ("abT", "cdT", "abcT", "cdbT", "badT", "ABT", "BCT") if rng.random() < 0.8}
        r = resolve_tf(ar, ai, br, bi, tags)
        lre = _rand_levels(rng, r["AB"], 3); lim = _rand_levels(rng, r["BC"], 2)
        cre, cim = QType(10, 4, True, 4, 0), QType(9, 3, True, 2, 2)
        plan = cgemul_plan(ar, ai, br, bi, tags, lre, lim, cre, cim, M, N, K, form=form)
        batch = 3
        wa = max(ar.word_bytes, ai.word_bytes); wb = max(br.word_bytes, bi.word_bytes)
        A = np.stack([rand_words(rng, ar, batch * M * K), rand_words(rng, ai, batch * M * K)], axis=1).astype(oracle.np_dtype(wa)).reshape(-1)
        B = np.stack([rand_words(rng, br, batch * K * N), rand_words(rng, bi, batch * K * N)], axis=1).astype(oracle.np_dtype(wb)).reshape(-1)
        want = oracle.cgemul_batched(plan, A, B, batch)
        Cd = torch.empty(batch * M * N * 2, dtype=torch.int32, device="cuda")
        ops.Qcgemul(Cd, dev(A), dev(B), plan, batch=batch)
        assert np.array_equal(Cd.cpu().numpy(), want), (M, N, K, form)


# ------------------------------------------------------------------------------------------------
# BASELINE configs at (or near) full size
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", range(7))
def test_c2_reduce_2p24_all_modes(mode):
    """Config 2: Qreduce of 2^24 x Qu<8,8>, C2a (8-entry list, last reused) and C2b (24-entry list)."""
    n = 1 << 24
    elem = QType(8, 8)
    x = torch.empty(n, dtype=torch.int32, device="cuda")
    ops.synth_fill(x, elem.raw_min, elem.raw_max, seed=1, stream_id=mode)
    xh = x.cpu().numpy()
    assert np.array_equal(xh[:4096], oracle.synth_fill(4096, 4, elem.raw_min, elem.raw_max, 1, mode))
    c2a = [QType(9 + l, 8 - l, True, mode, SAT_TCPL) for l in range(8)]
    c2b = c2a + [QType(17 + (l - 7 + 1) // 2, 0, True, mode, SAT_TCPL) for l in range(8, 24)]
    for lv, level in ((c2a, 9), (c2b, 17)):
        plan = reduce_plan(elem, lv, n)
        want, wparts = oracle.reduce_batched(plan, xh, 1, partial_level=level)
        out, ot, parts = ops.Qreduce(x, elem, lv, partial_level=level)
        assert np.array_equal(parts.cpu().numpy(), wparts)
        assert int(out.cpu()[0]) == int(want[0])
        assert ot == lv[-1]


@pytest.mark.parametrize("mode", range(7))
def test_c2_specialised_kernel_matches_oracle_and_generic(mode):
    """The compile-time-specialised Qreduce (no partials requested) against the oracle and against the
    runtime-descriptor kernel, batch of 3 vectors, both level lists."""
    n, batch = 1 << 24, 3
    elem = QType(8, 8)
    x = torch.empty(batch * n, dtype=torch.int32, device="cuda")
    ops.synth_fill(x, elem.raw_min, elem.raw_max, seed=2, stream_id=mode)
    c2a = [QType(9 + l, 8 - l, True, mode, SAT_TCPL) for l in range(8)]
    c2b = c2a + [QType(17 + (l - 7 + 1) // 2, 0, True, mode, SAT_TCPL) for l in range(8, 24)]
    for lv in (c2a, c2b):
        assert ops.plan_info_reduce(ops.make_reduce_plan(elem, lv, n)).specialized == 1
        out, _ = ops.Qreduce(x, elem, lv, length=n, batch=batch)
        gen, _ = ops.Qreduce(x, elem, lv, length=n, batch=batch, flags=abi.QB_F_NO_SPECIALIZE)
        assert torch.equal(out, gen)
        want, _ = oracle.reduce_batched(reduce_plan(elem, lv, n), x[n:2 * n].cpu().numpy(), 1)
        assert int(out.cpu()[1]) == int(want[0])


def test_c3_batched_64cubed_vs_oracle_and_shard_edges():
    """Config 3 at reduced batch against the oracle (first/last instances of a shard), full types."""
    a = QType(7, 8)
    mul = QType(14, 8, True, RND_CONV, SAT_TCPL)
    lv = [QType(15 + l, 8 - l, True, RND_POS_INF, SAT_TCPL) for l in range(6)]
    c = QType(17, 4, True, RND_POS_INF, SAT_TCPL)
    plan = ops.make_gemul_plan(a, a, c, 64, 64, 64, mul=mul, add=lv)
    batch = 96
    A = torch.empty(batch * 4096, dtype=torch.int32, device="cuda"); B = torch.empty_like(A)
    ops.synth_fill(A, a.raw_min, a.raw_max, 1, 0); ops.synth_fill(B, a.raw_min, a.raw_max, 1, 1)
    Cd = torch.empty(batch * 4096, dtype=torch.int32, device="cuda")
    ops.Qgemul(Cd, A, B, plan, batch=batch)
    want = oracle.gemul_batched(plan, A.cpu().numpy(), B.cpu().numpy(), batch)
    assert np.array_equal(Cd.cpu().numpy(), want)
    # the runtime-descriptor kernel and the specialised kernel must agree bit for bit
    C2 = torch.empty_like(Cd)
    ops.Qgemul(C2, A, B, plan, batch=batch, flags=abi.QB_F_NO_SPECIALIZE)
    assert torch.equal(Cd, C2)
    # saturation corner: min x min products (the only product the 14.8 MulArgs type cannot hold)
    A.fill_(a.raw_min); B.fill_(a.raw_min)
    ops.Qgemul(Cd, A, B, plan, batch=2)
    want = oracle.gemul_batched(plan, A[:8192].cpu().numpy(), B[:8192].cpu().numpy(), 2)
    assert np.array_equal(Cd[:8192].cpu().numpy(), want)


@pytest.mark.parametrize("seed", range(3))
def test_jit_specialised_kernels_vs_oracle(seed):
    """Plans outside the AOT registry get an NVRTC-specialised kernel at first use: Qgemul tiles,
    complex mat-vec and power-of-two Qreduce, each against the oracle and the runtime-descriptor kernel."""
    rng = np.random.default_rng(4000 + seed)
    # Qgemul
    for (M, N, K) in [(32, 32, 16), (64, 32, 64), (16, 64, 128)]:
        a = QType(int(rng.integers(3, 8)), int(rng.integers(3, 9)), True, 5, 0)
        b = QType(int(rng.integers(3, 8)), int(rng.integers(3, 9)), True, 5, 0)
        mul = QType(a.i + b.i - 1, max(a.f, b.f), True, int(rng.integers(0, 7)), int(rng.integers(0, 4)))
        lv = _rand_levels(rng, mul, int(rng.choice([1, 3, 7])))
        cc = QType(lv[-1].i, max(lv[-1].f - 1, 1 - lv[-1].i), True, int(rng.integers(0, 7)), int(rng.integers(0, 4)))
        plan = gemul_plan(a, b, mul, cc, lv, M, N, K)
        batch = 5
        A = rand_words(rng, a, batch * M * K); B = rand_words(rng, b, batch * K * N)
        want = oracle.gemul_batched(plan, A, B, batch)
        Cd = torch.empty(batch * M * N, dtype=torch.int32 if cc.word_bytes == 4 else torch.int64, device="cuda")
        try:
            ops.Qgemul(Cd, dev(A), dev(B), plan, batch=batch)
        except abi.QbError as e:
            assert e.code == abi.QB_EDOMAIN
            continue
        assert np.array_equal(Cd.cpu().numpy(), want), (M, N, K, a, b, mul, lv, cc)
        if cc.word_bytes == 4 and ops.plan_info_gemul(plan).arith_bits == 32:
            assert b"jit" in ops.plan_info_gemul(plan).kernel
    # Qreduce, power-of-two lengths
    for n in (1 << 16, 1 << 18, 1 << 20):
        elem = QType(int(rng.integers(4, 9)), int(rng.integers(3, 9)), True, int(rng.integers(0, 7)), int(rng.integers(0, 4)))
        lv = _rand_levels(rng, elem, int(rng.choice([2, 6, 20])))
        x = rand_words(rng, elem, n * 2)
        want, _ = oracle.reduce_batched(reduce_plan(elem, lv, n), x, 2)
        try:
            out, _ = ops.Qreduce(dev(x), elem, lv, length=n, batch=2)
            gen, _ = ops.Qreduce(dev(x), elem, lv, length=n, batch=2, flags=abi.QB_F_NO_SPECIALIZE)
        except abi.QbError as e:
            assert e.code == abi.QB_EDOMAIN
            continue
        assert np.array_equal(out.cpu().numpy(), want) and torch.equal(out, gen), (n, elem, lv)
    # complex mat-vec with random TF tags
    def rt():
        return QType(int(rng.integers(2, 6)), int(rng.integers(4, 10)), True, int(rng.integers(0, 7)), int(rng.integers(0, 4)))
    ar, br = rt(), rt()
    tags = {k: QType(int(rng.integers(6, 12)), int(rng.integers(4, 10)), True, int(rng.integers(0, 7)), int(rng.integers(0, 4)))
            for k in ("abT", "cdT", "abcT", "cdbT", "badT", "ABT", "BCT")}
    r = resolve_tf(ar, ar, br, br, tags)
    lre = _rand_levels(rng, r["AB"], 6); lim = _rand_levels(rng, r["BC"], 6)
    M, K, batch = 64, 32, 3
    plan = cgemul_plan(ar, ar, br, br, tags, lre, lim, QType(12, 3, True, 4, 0), QType(12, 3, True, 1, 2), M, 1, K, form=TREE_VARIADIC)
    A = np.stack([rand_words(rng, ar, batch * M * K), rand_words(rng, ar, batch * M * K)], axis=1).astype(np.int32).reshape(-1)
    B = np.stack([rand_words(rng, br, batch * K), rand_words(rng, br, batch * K)], axis=1).astype(np.int32).reshape(-1)
    want = oracle.cgemul_batched(plan, A, B, batch)
    Cd = torch.empty(batch * M * 2, dtype=torch.int32, device="cuda")
    try:
        ops.Qcgemul(Cd, dev(A), dev(B), plan, batch=batch)
        assert np.array_equal(Cd.cpu().numpy(), want)
    except abi.QbError as e:
        assert e.code == abi.QB_EDOMAIN


def test_c3_dynamic_clamp_elision_mixed_batch():
    """The tile kernel drops Qmul's saturation per instance when the instance's operand ranges allow it.
    A batch that mixes ordinary instances, instances where the two extreme operands meet in one product
    (min x min, the only saturating product of config 3) and instances where they exist but never meet
    must equal the oracle bit for bit."""
    a = QType(7, 8)
    mul = QType(14, 8, True, RND_CONV, SAT_TCPL)
    lv = [QType(15 + l, 8 - l, True, RND_POS_INF, SAT_TCPL) for l in range(6)]
    c = QType(17, 4, True, RND_POS_INF, SAT_TCPL)
    plan = ops.make_gemul_plan(a, a, c, 64, 64, 64, mul=mul, add=lv)
    batch = 6
    rng = np.random.default_rng(99)
    A = rng.integers(-32767, 32768, size=batch * 4096).astype(np.int32)
    B = rng.integers(-32767, 32768, size=batch * 4096).astype(np.int32)
    # instance 1: A[5,9] and B[9,3] are both the minimum -> product (5,3,k=9) saturates
    A[1 * 4096 + 5 + 9 * 64] = -32768; B[1 * 4096 + 9 + 3 * 64] = -32768
    # instance 3: minima present on different k -> never multiplied together
    A[3 * 4096 + 7 + 2 * 64] = -32768; B[3 * 4096 + 11 + 8 * 64] = -32768
    # instance 4: only A has the minimum
    A[4 * 4096 + 0] = -32768
    # instance 5: a whole column/row of minima
    A[5 * 4096 + 10 * 64: 5 * 4096 + 11 * 64] = -32768; B[5 * 4096 + 10: 5 * 4096 + 4096: 64] = -32768
    want = oracle.gemul_batched(plan, A, B, batch)
    Cd = torch.empty(batch * 4096, dtype=torch.int32, device="cuda")
    ops.Qgemul(Cd, dev(A), dev(B), plan, batch=batch)
    assert np.array_equal(Cd.cpu().numpy(), want)
    gen = torch.empty_like(Cd)
    ops.Qgemul(gen, dev(A), dev(B), plan, batch=batch, flags=abi.QB_F_NO_SPECIALIZE)
    assert torch.equal(Cd, gen)


def test_basic_complex_mul_gpu():
    """BasicComplexMul (4 mult / 2 add, the default complex product): element-wise goldens from the reference and a
    random complex Qgemul against the oracle."""
    from qublas_b200.qtypes import resolve_bcm
    for c in load_cases("ref_bcm.json"):
        xr, xi, yr, yi = (qt(c[k]) for k in ("x_re", "x_im", "y_re", "y_im"))
        wx = max(xr.word_bytes, xi.word_bytes); wy = max(yr.word_bytes, yi.word_bytes)
        x = np.array([i[0:2] for i in c["inputs"]], dtype=oracle.np_dtype(wx)).reshape(-1)
        y = np.array([i[2:4] for i in c["inputs"]], dtype=oracle.np_dtype(wy)).reshape(-1)
        out, _ = ops.tfcmul(dev(x), dev(y), xr, xi, yr, yi, tf_tags(c["tags"]), product="basic")
        got = out.cpu().numpy().reshape(-1, 2)
        assert [[int(a), int(b)] for a, b in got] == [o[0:2] for o in c["outputs"]], c["tags"]
    rng = np.random.default_rng(21)
    ar, ai, br, bi = QType(4, 7), QType(3, 8), QType(5, 6, True, 4, 2), QType(4, 6)
    tags = {"acT": QType(9, 8, True, 4, 0), "bdT": QType(8, 9, True, 0, 1), "adT": QType(9, 7), "bcT": QType(9, 8, True, 6, 3),
            "acbdT": QType(10, 6, True, 3, 0), "adbcT": QType(10, 7, True, 2, 2)}
    r = resolve_bcm(ar, ai, br, bi, tags)
    lre = _rand_levels(rng, r["acbd"], 4); lim = _rand_levels(rng, r["adbc"], 4)
    for (M, N, K, form) in [(4, 3, 16, TREE_VARIADIC), (8, 2, 7, TREE_VECTOR)]:
        plan = ops.make_cgemul_plan(ar, ai, br, bi, QType(12, 4, True, 4, 0), QType(12, 3), M, N, K, tf_tags=tags, add_re=lre, add_im=lim,
                                    form=form, product="basic")
        batch = 4
        A = np.stack([rand_words(rng, ar, batch * M * K), rand_words(rng, ai, batch * M * K)], axis=1).astype(np.int32).reshape(-1)
        B = np.stack([rand_words(rng, br, batch * K * N), rand_words(rng, bi, batch * K * N)], axis=1).astype(np.int32).reshape(-1)
        want = oracle.cgemul_batched(plan, A, B, batch)
        Cd = torch.empty(batch * M * N * 2, dtype=torch.int32, device="cuda")
        ops.Qcgemul(Cd, dev(A), dev(B), plan, batch=batch)
        assert np.array_equal(Cd.cpu().numpy(), want), (M, N, K, form)


def test_real_complex_products_gpu():
    """real (x) complex Qmul and mixed real/complex Qgemul (next row 8(f).2): the reference's own answers
    (tests/golden/ref_rc.json) element-wise and as 1xK dot products, then random shapes against the oracle."""
    from qublas_b200.qtypes import resolve_rc
    from tests.helpers import rc_targs
    n = 0
    for c in load_cases("ref_rc.json"):
        if c["kind"] == "rc" and c["op"] == "mul":
            r, cre, cim = qt(c["r"]), qt(c["c_re"]), qt(c["c_im"])
            real = np.array([i[0] for i in c["inputs"]], dtype=np.int32)
            cplx = np.array([i[1:3] for i in c["inputs"]], dtype=np.int32).reshape(-1)
            if c["real_first"]:
                out, _ = ops.tfcmul(dev(real), dev(cplx), r, None, cre, cim, rc_targs(c["targs"]), product="real_a")
            else:
                out, _ = ops.tfcmul(dev(cplx), dev(real), cre, cim, r, None, rc_targs(c["targs"]), product="real_b")
            got = out.cpu().numpy().reshape(-1, 2)
            assert [[int(a), int(b)] for a, b in got] == [o[0:2] for o in c["outputs"]], c["targs"]
            n += len(c["inputs"])
        elif c["kind"] == "rcgemul":
            K = c["K"]; rf = c["real_first"]
            r, cre, cim = qt(c["r"]), qt(c["c_re"]), qt(c["c_im"])
            rr = resolve_rc(rf, r, cre, cim, rc_targs(c["targs"]))
            lv = [qt(t) for t in c["levels"]]
            o_re = tree_out_type(K, rr["ac"], lv, TREE_VARIADIC); o_im = tree_out_type(K, rr["ad" if rf else "bc"], lv, TREE_VARIADIC)
            args = (r, None, cre, cim) if rf else (cre, cim, r, None)
            plan = ops.make_cgemul_plan(*args, o_re, o_im, 1, 1, K, tf_tags=rc_targs(c["targs"]), add_re=lv, form=TREE_VARIADIC,
                                        product="real_a" if rf else "real_b")
            batch = len(c["inputs"])
            v = np.array(c["inputs"], dtype=np.int32).reshape(batch, K, 3)
            real = np.ascontiguousarray(v[:, :, 0]).reshape(-1); cplx = np.ascontiguousarray(v[:, :, 1:3]).reshape(-1)
            Cd = torch.empty(batch * 2, dtype=torch.int32, device="cuda"); assert max(o_re.word_bytes, o_im.word_bytes) == 4
            ops.Qcgemul(Cd, dev(real if rf else cplx), dev(cplx if rf else real), plan, batch=batch)
            got = Cd.cpu().numpy().reshape(batch, 2)
            assert [[int(a), int(b)] for a, b in got] == [o[0:2] for o in c["outputs"]], (K, rf)
            n += batch
    assert n >= 448
    rng = np.random.default_rng(33)
    for (M, N, K, form, rf, ta, tb, wide) in [(4, 3, 16, TREE_VARIADIC, True, False, False, False), (8, 2, 7, TREE_VECTOR, False, True, False, False),
                                              (3, 5, 9, TREE_VECTOR, True, False, True, False), (6, 4, 33, TREE_VARIADIC, False, False, False, True)]:
        r, cre, cim = QType(5, 7), QType(4, 8, True, 2, 1), QType(3, 9)
        tags = {"realT": QType(22 if wide else 9, 14 if wide else 8, True, 4, 0), "imagT": QType(8, 9, True, 3, 2)}
        rr = resolve_rc(rf, r, cre, cim, tags)
        lre = _rand_levels(rng, rr["ac"], 4); lim = _rand_levels(rng, rr["ad" if rf else "bc"], 4)
        args = (r, None, cre, cim) if rf else (cre, cim, r, None)
        plan = ops.make_cgemul_plan(*args, QType(12, 4, True, 4, 0), QType(12, 3), M, N, K, tf_tags=tags, add_re=lre, add_im=lim,
                                    form=form, product="real_a" if rf else "real_b", trans_a=ta, trans_b=tb)
        batch = 3
        ea, eb = batch * M * K, batch * K * N
        def cplx_words(n):
            return np.stack([rand_words(rng, cre, n), rand_words(rng, cim, n)], axis=1).astype(np.int32).reshape(-1)
        A = rand_words(rng, r, ea).astype(np.int32) if rf else cplx_words(ea)
        B = cplx_words(eb) if rf else rand_words(rng, r, eb).astype(np.int32)
        want = oracle.cgemul_batched(plan, A, B, batch)
        Cd = torch.empty(batch * M * N * 2, dtype=torch.int32, device="cuda")
        ops.Qcgemul(Cd, dev(A), dev(B), plan, batch=batch)
        assert np.array_equal(Cd.cpu().numpy(), want), (M, N, K, form, rf)
        Ch = np.zeros(batch * M * N * 2, dtype=np.int32)                 # host buffers through the staging pipeline
        ops.Qcgemul(Ch, A, B, plan, batch=batch)
        assert np.array_equal(Ch, want)


def test_elementwise_tensor_ops_vs_reference_goldens_and_oracle():
    """Element-wise Qmul/Qadd/Qsub on tensors (next row 8(f).3): every scalar golden of the reference, run as
    one-element-per-input tensors on the GPU, plus random large tensors against the oracle."""
    n = 0
    for c in load_cases("ref_scalar.json"):
        if c["op"] == "conv":
            continue
        a, b = qt(c["a"]), qt(c["b"])
        to = qt(c["to"]) if "to" in c else c.get("tags")
        A = np_words([i[0] for i in c["inputs"]], a); B = np_words([i[1] for i in c["inputs"]], b)
        out, rt = ops.elementwise(c["op"], dev(A), dev(B), a, b, to)
        assert [int(v) for v in out.cpu().numpy()] == [o[0] for o in c["outputs"]], c
        assert [rt.i, rt.f, int(rt.s), rt.q, rt.o] == c["outputs"][0][1:6]
        n += len(A)
    assert n > 8000
    rng = np.random.default_rng(8)
    for op in ("mul", "add", "sub"):
        a = QType(9, 14, True, 3, 1); b = QType(12, 9, True, 2, 2)             # 64-bit register class for the product
        to = QType(15, 10, True, 4, 0)
        A = rand_words(rng, a, 100003); B = rand_words(rng, b, 100003)
        out, _ = ops.elementwise(op, dev(A), dev(B), a, b, to)
        want = oracle.elementwise(ops.make_ew_plan(op, a, b, 100003, to), A, B)
        assert np.array_equal(out.cpu().numpy(), want), op


def test_wide_two_word_slice_gpu():
    """Intermediates of 65..126 bits (products of two > 32-bit operands, wide aligned sums) run in __int128 inside the
    64-bit kernels: reference goldens through the element-wise entry point, the fenced cases are refused, and a
    random Qgemul with 82-bit products against the oracle."""
    n = refused = 0
    for c in load_cases("ref_wide.json"):
        a, b, to = qt(c["a"]), qt(c["b"]), qt(c["to"])
        A = np_words([i[0] for i in c["inputs"]], a); B = np_words([i[1] for i in c["inputs"]], b)
        try:
            out, _ = ops.elementwise(c["op"], dev(A), dev(B), a, b, to)
        except abi.QbError as e:
            assert e.code == abi.QB_EDOMAIN and c["band"], (c["a"], c["b"], c["to"], str(e))
            refused += 1
            continue
        # whatever the fence admits (incl. `band` candidates whose true range stays below 2^63) must equal the reference
        assert [int(v) for v in out.cpu().numpy()] == [o[0] for o in c["outputs"]], (c["op"], c["a"], c["b"], c["to"])
        n += len(A)
    assert n >= 48 * 40 and refused >= 12
    rng = np.random.default_rng(77)
    a = QType(20, 20, True, 5, 0); b = QType(22, 18, True, 5, 0)
    mul = QType(24, 20, True, 3, 0)                            # 83-bit product >> 18 with RND::INF
    lv = [QType(25 + l, 20 - 2 * l, True, (l * 2) % 4, l % 3) for l in range(4)]
    cc = QType(30, 10, True, 1, 2)
    M, N, K, batch = 8, 6, 16, 3
    plan = gemul_plan(a, b, mul, cc, lv, M, N, K)
    info = ops.plan_info_gemul(plan)
    assert info.arith_bits == 64 and info.max_inter_bits > 64
    A = rand_words(rng, a, batch * M * K); B = rand_words(rng, b, batch * K * N)
    want = oracle.gemul_batched(plan, A, B, batch)
    Cd = torch.empty(batch * M * N, dtype=torch.int64, device="cuda")
    ops.Qgemul(Cd, dev(A), dev(B), plan, batch=batch)
    assert np.array_equal(Cd.cpu().numpy(), want)


def test_synth_fill_matches_oracle_all_word_sizes():
    for wb, lo, hi in [(1, -127, 127), (2, -(1 << 15), (1 << 15) - 1), (4, -65536, 65535), (8, -(1 << 40), (1 << 40))]:
        t = torch.empty(10007, dtype={1: torch.int8, 2: torch.int16, 4: torch.int32, 8: torch.int64}[wb], device="cuda")
        ops.synth_fill(t, lo, hi, seed=3, stream_id=9, first_index=12345)
        assert np.array_equal(t.cpu().numpy(), oracle.synth_fill(10007, wb, lo, hi, 3, 9, 12345))
        assert int(t.min()) >= lo and int(t.max()) <= hi


def test_launch_counter_counts_kernels():
    lib = abi.load()
    lib.qb200_launch_count_reset()
    x = torch.zeros(1 << 16, dtype=torch.int32, device="cuda")
    ops.Qreduce(x, QType(8, 8), [QType(9, 8)])
    assert lib.qb200_launch_count() == 2          # 2^16 -> 16 partials -> root


def test_two_devices_in_one_process():
    """The C ABI takes a device ordinal per call: function attributes, occupancy figures and JIT-built modules are kept
    per device, so one process can drive several GPUs (needs >= 2 visible GPUs; `gpurun --gpus 2`)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rng = np.random.default_rng(77)
    a = QType(7, 8); mul = QType(14, 8, True, RND_CONV, SAT_TCPL)
    lv = [QType(15 + l, 8 - l, True, RND_POS_INF, SAT_TCPL) for l in range(6)]; c = QType(17, 4, True, RND_POS_INF, SAT_TCPL)
    aot = ops.make_gemul_plan(a, a, c, 64, 64, 64, mul=mul, add=lv)                        # ahead-of-time kernel (dynamic smem attribute)
    ja = QType(5, 6); jmul = QType(9, 6, True, 4, 0); jlv = [QType(10 + l, 6 - l, True, 2, 0) for l in range(4)]
    jit = gemul_plan(ja, ja, jmul, QType(13, 2, True, 0, 0), jlv, 32, 32, 16)              # NVRTC-built kernel (module per context)
    i8 = QType(7, 0); acc = QType(30, 0)
    tc = ops.make_gemul_plan(i8, i8, acc, 128, 256, 128, mul=QType(15, 0), add=[acc])      # tcgen05 path
    for plan, t, batch in ((aot, a, 6), (jit, ja, 5), (tc, i8, 1)):
        M, N, K = plan.M, plan.N, plan.K
        A = rand_words(rng, t, batch * M * K).astype(np.int32); B = rand_words(rng, t, batch * K * N).astype(np.int32)
        want = oracle.gemul_batched(plan, A, B, batch)
        for d in (0, 1, 0):
            dv = torch.device("cuda", d)
            Cd = torch.empty(batch * M * N, dtype=torch.int32, device=dv)
            ops.Qgemul(Cd, torch.from_numpy(A).to(dv), torch.from_numpy(B).to(dv), plan, batch=batch)
            assert np.array_equal(Cd.cpu().numpy(), want), (d, M, N, K)


@pytest.mark.parametrize("seed", range(3))
def test_jit_thread_kernel_odd_shapes_vs_oracle(seed):
    """Shapes the register-tile template does not take (K not a power of two, M/N not multiples of 4, tiny instances,
    mat-vec, transposed operands, the 64-bit register class, both tree forms) get the NVRTC-specialised
    thread-per-output kernel: bit-exact against the oracle and against the runtime-descriptor kernel."""
    rng = np.random.default_rng(9100 + seed)
    shapes = [(48, 48, 48), (20, 12, 100), (16, 16, 16), (64, 1, 64), (3, 5, 7), (2, 2, 1), (33, 7, 65), (8, 8, 200)]
    n_jit = 0
    for (M, N, K) in shapes:
        form = TREE_VECTOR if rng.random() < 0.5 else TREE_VARIADIC
        ta, tb = bool(rng.random() < 0.3), bool(rng.random() < 0.3)
        wide = rng.random() < 0.25
        a = QType(int(rng.integers(3, 8)) + (14 if wide else 0), int(rng.integers(3, 9)), True, 5, 0)
        b = QType(int(rng.integers(3, 8)) + (14 if wide else 0), int(rng.integers(3, 9)), True, 5, 0)
        mul = QType(a.i + b.i - 1, max(a.f, b.f), True, int(rng.integers(0, 7)), int(rng.integers(0, 4)))
        lv = _rand_levels(rng, mul, int(rng.choice([1, 3, 7])))
        cc = QType(lv[-1].i, max(lv[-1].f - 1, 1 - lv[-1].i), True, int(rng.integers(0, 7)), int(rng.integers(0, 4)))
        plan = gemul_plan(a, b, mul, cc, lv, M, N, K, trans_a=ta, trans_b=tb, form=form)
        batch = 7
        A = rand_words(rng, a, batch * M * K); B = rand_words(rng, b, batch * K * N)
        tdt = torch.int32 if cc.word_bytes == 4 else torch.int64
        Cd = torch.empty(batch * M * N, dtype=tdt, device="cuda")
        try:
            ops.Qgemul(Cd, dev(A), dev(B), plan, batch=batch)
        except abi.QbError as e:
            assert e.code == abi.QB_EDOMAIN
            continue
        want = oracle.gemul_batched(plan, A, B, batch)
        assert np.array_equal(Cd.cpu().numpy(), want), (M, N, K, form, ta, tb, a, b, mul, lv, cc)
        gen = torch.empty_like(Cd)
        ops.Qgemul(gen, dev(A), dev(B), plan, batch=batch, flags=abi.QB_F_NO_SPECIALIZE)
        assert torch.equal(Cd, gen)
        if ops.plan_info_gemul(plan).kernel == b"k_gemul_thread<jit>":
            n_jit += 1
    assert n_jit >= 5


@pytest.mark.parametrize("kind", ["tf", "basic", "real_a", "real_b"])
def test_jit_complex_thread_kernel_all_product_kinds(kind):
    """Complex Qgemul outside the mat-vec template (all four product kinds, ragged shapes, both forms, transposes)
    runs the NVRTC-specialised thread-per-output kernel: bit-exact against the oracle and the runtime-descriptor kernel."""
    from qublas_b200.qtypes import resolve_bcm, resolve_rc
    rng = np.random.default_rng({"tf": 1, "basic": 2, "real_a": 3, "real_b": 4}[kind] + 500)
    n_jit = 0
    for (M, N, K) in [(5, 3, 12), (8, 8, 40), (16, 2, 7), (4, 4, 64), (33, 1, 100), (2, 2, 1)]:
        form = TREE_VECTOR if rng.random() < 0.5 else TREE_VARIADIC
        ta, tb = bool(rng.random() < 0.3), bool(rng.random() < 0.3)
        def rt(lo=2, hi=7):
            return QType(int(rng.integers(lo, hi)), int(rng.integers(1, 8)), True, int(rng.integers(0, 7)), int(rng.integers(0, 4)))
        ar, ai, br, bi = rt(), rt(), rt(), rt()
        if kind == "tf":
            tags = {k: rt(5, 12) for k in ("abT", "cdT", "abcT", "cdbT", "badT", "ABT", "BCT") if rng.random() < 0.8}
            r = resolve_tf(ar, ai, br, bi, tags); leaf = (r["AB"], r["BC"]); args = (ar, ai, br, bi)
        elif kind == "basic":
            tags = {k: rt(5, 12) for k in ("acT", "bdT", "adT", "bcT", "acbdT", "adbcT") if rng.random() < 0.8}
            r = resolve_bcm(ar, ai, br, bi, tags); leaf = (r["acbd"], r["adbc"]); args = (ar, ai, br, bi)
        elif kind == "real_a":
            tags = {"realT": rt(6, 12), "imagT": rt(6, 12)}
            r = resolve_rc(True, ar, br, bi, tags); leaf = (r["ac"], r["ad"]); args = (ar, None, br, bi)
        else:
            tags = {"realT": rt(6, 12), "imagT": rt(6, 12)}
            r = resolve_rc(False, br, ar, ai, tags); leaf = (r["ac"], r["bc"]); args = (ar, ai, br, None)
        lre = _rand_levels(rng, leaf[0], 3); lim = _rand_levels(rng, leaf[1], 2)
        plan = ops.make_cgemul_plan(*args, QType(10, 4, True, 4, 0), QType(9, 3, True, 2, 2), M, N, K, tf_tags=tags, add_re=lre, add_im=lim,
                                    form=form, product=kind, trans_a=ta, trans_b=tb)
        batch = 5
        def words(re_t, im_t, n):
            if im_t is None:
                return rand_words(rng, re_t, n).astype(np.int32)
            return np.stack([rand_words(rng, re_t, n), rand_words(rng, im_t, n)], axis=1).astype(np.int32).reshape(-1)
        A = words(args[0], args[1], batch * M * K); B = words(args[2], args[3], batch * K * N)
        try:
            Cd = torch.empty(batch * M * N * 2, dtype=torch.int32, device="cuda")
            ops.Qcgemul(Cd, dev(A), dev(B), plan, batch=batch)
        except abi.QbError as e:
            assert e.code == abi.QB_EDOMAIN
            continue
        want = oracle.cgemul_batched(plan, A, B, batch)
        assert np.array_equal(Cd.cpu().numpy(), want), (kind, M, N, K, form, ta, tb)
        gen = torch.empty_like(Cd)
        ops.Qcgemul(gen, dev(A), dev(B), plan, batch=batch, flags=abi.QB_F_NO_SPECIALIZE)
        assert torch.equal(Cd, gen)
        n_jit += ops.plan_info_cgemul(plan).kernel == b"k_cgemul_thread<jit>"
    assert n_jit >= 4


@pytest.mark.parametrize("seed", range(2))
def test_jit_any_length_reduce_vs_oracle(seed):
    """Vector-form Qreduce at lengths that are not 2^16..2^24 powers of two: the NVRTC-specialised any-length kernel
    (complete chunks on the fast path, the ragged last chunk with presence tracking and carry conversions) against the
    oracle and the runtime-descriptor kernels; power-of-two variadic plans are routed to the same kernels."""
    rng = np.random.default_rng(700 + seed)
    n_jit = 0
    for n in (64, 65, 100, 255, 256, 257, 1000, 4095, 4096, 4097, 65535, 100000, (1 << 20) + 1, 3 << 18):
        elem = QType(int(rng.integers(4, 9)), int(rng.integers(3, 9)), True, int(rng.integers(0, 7)), int(rng.integers(0, 4)))
        lv = _rand_levels(rng, elem, int(rng.choice([2, 6, 20])))
        batch = 3 if n < 200000 else 2
        x = rand_words(rng, elem, batch * n).astype(np.int32)
        plan = reduce_plan(elem, lv, n)
        try:
            out, _ = ops.Qreduce(dev(x), elem, lv, length=n, batch=batch)
        except abi.QbError as e:
            assert e.code == abi.QB_EDOMAIN
            continue
        want, _ = oracle.reduce_batched(plan, x, batch)
        assert np.array_equal(out.cpu().numpy(), want), (n, elem, lv)
        gen, _ = ops.Qreduce(dev(x), elem, lv, length=n, batch=batch, flags=abi.QB_F_NO_SPECIALIZE)
        assert torch.equal(out, gen)
        if out.dtype == torch.int32 and ops.plan_info_reduce(plan).arith_bits == 32:
            assert ops.plan_info_reduce(plan).kernel == b"k_reduce_any<jit>", n
            n_jit += 1
    assert n_jit >= 8
    # power of two, variadic form == vector form
    elem = QType(8, 8); lv = [QType(9 + l, 8 - l, True, 0, 0) for l in range(8)]
    n = 1 << 12
    x = rand_words(rng, elem, 2 * n).astype(np.int32)
    out, _ = ops.Qreduce(dev(x), elem, lv, length=n, batch=2, form=TREE_VARIADIC)
    want, _ = oracle.reduce_batched(reduce_plan(elem, lv, n, form=TREE_VARIADIC), x, 2)
    assert np.array_equal(out.cpu().numpy(), want)


@pytest.mark.parametrize("n,batch", [(513, 7), (1000, 5), (65535, 3), (65537, 3), (100000, 4), (3 << 20, 2), ((1 << 24) - 1, 1), (256 * 257 + 255, 2)])
def test_variadic_qreduce_of_lengths_that_are_not_a_power_of_two(n, batch):
    """the variadic form (QuBLAS.h:4942-4949: the odd node of a level is added after the levels above
    it) on long vectors: whole 256-node chunks are reduced pass by pass by the warp kernels, k_reduce_var_finish streams
    what is left through the level-indexed stack.  Bit-exact against the oracle and the one-thread-per-vector kernel."""
    rng = np.random.default_rng(n)
    for mode in (RND_POS_INF, RND_CONV):
        elem = QType(8, 8)
        lv = [QType(9 + l, 8 - l, True, mode, SAT_TCPL) for l in range(8)] + [QType(17 + (l - 6) // 2, 0, True, mode, SAT_TCPL) for l in range(8, 24)]
        plan = reduce_plan(elem, lv, n, form=TREE_VARIADIC)
        x = rand_words(rng, elem, batch * n).astype(np.int32)
        out, _ = ops.Qreduce(dev(x), elem, lv, length=n, batch=batch, form=TREE_VARIADIC)
        assert ops.plan_info_reduce(plan).kernel == b"k_reduce_var<jit>", ops.plan_info_reduce(plan).kernel
        want, _ = oracle.reduce_batched(plan, x, batch)
        assert np.array_equal(out.cpu().numpy(), want), (n, mode)
        if n <= 100000:
            gen, _ = ops.Qreduce(dev(x), elem, lv, length=n, batch=batch, form=TREE_VARIADIC, flags=abi.QB_F_NO_SPECIALIZE)
            assert torch.equal(out, gen)
            # a buffer that does not start on a 16-byte boundary stays on the one-thread-per-vector kernel: same bits
            buf = dev(np.concatenate([np.zeros(1, np.int32), x]))
            out2, _ = ops.Qreduce(buf[1:], elem, lv, length=n, batch=batch, form=TREE_VARIADIC)
            assert np.array_equal(out2.cpu().numpy(), want), (n, mode, "unaligned")


def test_jit_elementwise_kernels_vs_runtime_descriptor_and_oracle():
    """The NVRTC-specialised element-wise kernels (complex products of all four kinds, real Qmul/Qadd/Qsub, converting
    assignment) on odd lengths (ragged tail of the 4-way unrolled loop), 32- and 64-bit register classes: identical to the
    runtime-descriptor kernels, and to the oracle on a prefix."""
    from qublas_b200.qtypes import resolve_bcm, resolve_rc
    rng = np.random.default_rng(4242)
    n = 100003
    for wide in (False, True):
        t = QType(7, 8) if not wide else QType(20, 10)
        to = QType(10, 6, True, 4, 0) if not wide else QType(36, 9, True, 3, 2)
        a = rand_words(rng, t, n).astype(np.int32); b = rand_words(rng, t, n).astype(np.int32)
        for op in ("mul", "add", "sub"):
            out, rt = ops.elementwise(op, dev(a), dev(b), t, t, to=to)
            gen, _ = ops.elementwise(op, dev(a), dev(b), t, t, to=to, flags=abi.QB_F_NO_SPECIALIZE)
            assert torch.equal(out, gen), (op, wide)
            want = oracle.elementwise(ops.make_ew_plan(op, t, t, 257, to), a[:257], b[:257])
            assert np.array_equal(out[:257].cpu().numpy(), want), (op, wide)
        cv = ops.convert(dev(a), t, to)
        assert torch.equal(cv, ops.convert(dev(a), t, to, flags=abi.QB_F_NO_SPECIALIZE))
    c = QType(3, 12); P = QType(6, 12); S = QType(7, 12)
    x = rand_words(rng, c, 2 * n).astype(np.int32); y = rand_words(rng, c, 2 * n).astype(np.int32); r1 = rand_words(rng, c, n).astype(np.int32)
    tf = {"abT": QType(4, 12), "cdT": QType(4, 11), "abcT": P, "cdbT": P, "badT": P, "ABT": S, "BCT": S}
    bc = {"acT": P, "bdT": QType(6, 11, True, 4, 1), "adT": P, "bcT": P, "acbdT": S, "adbcT": QType(7, 10, True, 0, 2)}
    rc = {"realT": P, "imagT": QType(6, 10, True, 2, 0)}
    for product, tags, xx, yy, ty in (("tf", tf, x, y, (c, c, c, c)), ("basic", bc, x, y, (c, c, c, c)), ("real_a", rc, r1, y, (c, None, c, c)),
                                      ("real_b", rc, x, r1, (c, c, c, None))):
        out, _ = ops.tfcmul(dev(xx), dev(yy), *ty, tags, product=product)
        gen, _ = ops.tfcmul(dev(xx), dev(yy), *ty, tags, product=product, flags=abi.QB_F_NO_SPECIALIZE)
        assert torch.equal(out, gen), product


@pytest.mark.parametrize("form", [TREE_VECTOR, TREE_VARIADIC])
def test_jit_tile_kernel_k_multiple_of_8_vs_oracle(form):
    """The register-tile template with K = 8 x (not a power of two): pending nodes are combined by the compile-time
    flush program of the requested tree form."""
    rng = np.random.default_rng(31 + form)
    for (M, N, K) in [(64, 64, 48), (32, 32, 24), (64, 32, 40), (64, 64, 56), (32, 64, 104), (48, 48, 48), (40, 40, 16), (36, 60, 24), (28, 28, 8)]:
        a = QType(int(rng.integers(3, 8)), int(rng.integers(3, 9)), True, 5, 0)
        b = QType(int(rng.integers(3, 8)), int(rng.integers(3, 9)), True, 5, 0)
        mul = QType(a.i + b.i - 1, max(a.f, b.f), True, int(rng.integers(0, 7)), int(rng.integers(0, 4)))
        lv = _rand_levels(rng, mul, int(rng.choice([1, 3, 7])))
        cc = QType(lv[-1].i, max(lv[-1].f - 1, 1 - lv[-1].i), True, int(rng.integers(0, 7)), int(rng.integers(0, 4)))
        ta, tb = bool(rng.random() < 0.4), bool(rng.random() < 0.4)              # transposed operands use the same template
        plan = gemul_plan(a, b, mul, cc, lv, M, N, K, form=form, trans_a=ta, trans_b=tb)
        batch = 5
        A = rand_words(rng, a, batch * M * K); B = rand_words(rng, b, batch * K * N)
        Cd = torch.empty(batch * M * N, dtype=torch.int32 if cc.word_bytes == 4 else torch.int64, device="cuda")
        try:
            ops.Qgemul(Cd, dev(A), dev(B), plan, batch=batch)
        except abi.QbError as e:
            assert e.code == abi.QB_EDOMAIN
            continue
        want = oracle.gemul_batched(plan, A, B, batch)
        assert np.array_equal(Cd.cpu().numpy(), want), (M, N, K, form, ta, tb, a, b, mul, lv, cc)
        if cc.word_bytes == 4 and ops.plan_info_gemul(plan).arith_bits == 32:
            assert ops.plan_info_gemul(plan).kernel == b"k_gemul_tile<jit>", (M, N, K)


@pytest.mark.parametrize("form", [TREE_VECTOR, TREE_VARIADIC])
def test_jit_tile_kernel_small_instances_in_groups(form):
    """Instances with fewer than 48 tile threads: one CTA of the register-tile template works on a group of consecutive
    instances at a time (two bulk copies per group); batches that are not a multiple of the group exercise the short
    last group."""
    rng = np.random.default_rng(77 + form)
    for (M, N, K, batch) in [(16, 16, 16, 37), (8, 8, 64, 100), (20, 12, 64, 35), (4, 4, 8, 513), (12, 8, 24, 64), (16, 16, 16, 3)]:
        a = QType(int(rng.integers(3, 8)), int(rng.integers(3, 9)), True, 5, 0)
        b = QType(int(rng.integers(3, 8)), int(rng.integers(3, 9)), True, 5, 0)
        mul = QType(a.i + b.i - 1, max(a.f, b.f), True, int(rng.integers(0, 7)), int(rng.integers(0, 4)))
        lv = _rand_levels(rng, mul, int(rng.choice([1, 3, 7])))
        cc = QType(lv[-1].i, max(lv[-1].f - 1, 1 - lv[-1].i), True, int(rng.integers(0, 7)), int(rng.integers(0, 4)))
        ta, tb = bool(rng.random() < 0.4), bool(rng.random() < 0.4)
        plan = gemul_plan(a, b, mul, cc, lv, M, N, K, form=form, trans_a=ta, trans_b=tb)
        A = rand_words(rng, a, batch * M * K); B = rand_words(rng, b, batch * K * N)
        Cd = torch.empty(batch * M * N, dtype=torch.int32 if cc.word_bytes == 4 else torch.int64, device="cuda")
        try:
            ops.Qgemul(Cd, dev(A), dev(B), plan, batch=batch)
        except abi.QbError as e:
            assert e.code == abi.QB_EDOMAIN
            continue
        want = oracle.gemul_batched(plan, A, B, batch)
        assert np.array_equal(Cd.cpu().numpy(), want), (M, N, K, batch, form, a, b, mul, lv, cc)
        if cc.word_bytes == 4 and ops.plan_info_gemul(plan).arith_bits == 32:
            assert ops.plan_info_gemul(plan).kernel == b"k_gemul_tile<jit>", (M, N, K)


def test_batch_strides_on_the_specialised_kernels():
    """Batch strides (0 = an operand shared by every instance, larger than dense = padded layouts) are part of the plan and
    are baked into the run-time specialised kernels: real tile and thread templates, complex rows and thread templates."""
    rng = np.random.default_rng(515)
    a = QType(7, 8); mul = QType(14, 8, True, RND_CONV, SAT_TCPL)
    lv = [QType(15 + l, 8 - l, True, RND_POS_INF, SAT_TCPL) for l in range(6)]; c = QType(17, 4, True, RND_POS_INF, SAT_TCPL)
    for (M, N, K, sa, sb, sc, kernel) in [(64, 64, 64, 0, 4096, 4096, b"k_gemul_tile<jit>"), (64, 64, 64, 4096, 0, 4100, b"k_gemul_tile<jit>"),
                                          (32, 32, 16, 512 + 8, 512 + 4, 1024 + 12, b"k_gemul_tile<jit>"), (5, 7, 12, 0, 84 + 3, 35 + 1, b"k_gemul_thread<jit>")]:
        plan = ops.make_gemul_plan(a, a, c, M, N, K, mul=mul, add=lv)
        plan.stride_a, plan.stride_b, plan.stride_c = sa, sb, sc
        batch = 6
        A = rand_words(rng, a, (batch - 1) * sa + M * K).astype(np.int32); B = rand_words(rng, a, (batch - 1) * sb + K * N).astype(np.int32)
        want = oracle.gemul_batched(plan, A, B, batch)
        Cd = torch.zeros((batch - 1) * sc + M * N, dtype=torch.int32, device="cuda")
        ops.Qgemul(Cd, dev(A), dev(B), plan, batch=batch)
        got = Cd.cpu().numpy()
        for i in range(batch):
            assert np.array_equal(got[i * sc:i * sc + M * N], want[i * sc:i * sc + M * N]), (M, N, K, sa, sb, sc, i)
        assert ops.plan_info_gemul(plan).kernel == kernel
    # complex: a channel matrix shared by all instances (stride 0), mat-vec rows template and thread template
    t = QType(3, 12); P = QType(6, 12); S = QType(7, 12)
    tags = {"abT": QType(4, 12), "cdT": QType(4, 11), "abcT": P, "cdbT": P, "badT": P, "ABT": S, "BCT": S}
    clv = [QType(8 + l, 12 - l, True, RND_CONV, SAT_TCPL) for l in range(6)]
    for (M, N, K, kernel) in [(64, 1, 64, b"k_cgemv_rows<jit>"), (6, 2, 10, b"k_cgemul_thread<jit>")]:
        plan = ops.make_cgemul_plan(t, t, t, t, clv[-1], clv[-1], M, N, K, tf_tags=tags, add_re=clv)
        plan.stride_a = 0
        batch = 9
        H = np.stack([rand_words(rng, t, M * K), rand_words(rng, t, M * K)], axis=1).astype(np.int32).reshape(-1)
        X = np.stack([rand_words(rng, t, batch * K * N), rand_words(rng, t, batch * K * N)], axis=1).astype(np.int32).reshape(-1)
        want = oracle.cgemul_batched(plan, H, X, batch)
        Y = torch.empty(batch * M * N * 2, dtype=torch.int32, device="cuda")
        ops.Qcgemul(Y, dev(H), dev(X), plan, batch=batch)
        assert np.array_equal(Y.cpu().numpy(), want), (M, N, K)
        assert ops.plan_info_cgemul(plan).kernel == kernel


def test_packed_int16_operand_words_on_the_tile_template():
    """qb_qdesc.word_bytes = 2: operands arrive as packed int16 words (half the HBM / PCIe bytes), are staged as they are
    and widened by the conversion pass.  Same results as the oracle on the same packed buffers, and as the int32 call."""
    rng = np.random.default_rng(1616)
    a = QType(7, 8); mul = QType(14, 8, True, RND_CONV, SAT_TCPL)
    lv = [QType(15 + l, 8 - l, True, RND_POS_INF, SAT_TCPL) for l in range(6)]; c = QType(17, 4, True, RND_POS_INF, SAT_TCPL)
    for (M, N, K, batch, ta) in [(64, 64, 64, 9, False), (32, 32, 24, 7, True), (16, 16, 16, 37, False), (64, 32, 40, 5, False)]:
        p16 = ops.make_gemul_plan(a, a, c, M, N, K, mul=mul, add=lv, word_bytes=(2, 2, 0), trans_a=ta)
        p32 = ops.make_gemul_plan(a, a, c, M, N, K, mul=mul, add=lv, trans_a=ta)
        A = rng.integers(a.raw_min, a.raw_max + 1, size=batch * M * K).astype(np.int16)
        B = rng.integers(a.raw_min, a.raw_max + 1, size=batch * K * N).astype(np.int16)
        A[:5] = a.raw_min; B[:5] = a.raw_min
        want = oracle.gemul_batched(p16, A, B, batch)
        C16 = torch.empty(batch * M * N, dtype=torch.int32, device="cuda")
        ops.Qgemul(C16, dev(A), dev(B), p16, batch=batch)
        assert np.array_equal(C16.cpu().numpy(), want), (M, N, K)
        C32 = torch.empty_like(C16)
        ops.Qgemul(C32, dev(A.astype(np.int32)), dev(B.astype(np.int32)), p32, batch=batch)
        assert torch.equal(C16, C32)
        assert ops.plan_info_gemul(p16).kernel == b"k_gemul_tile<jit>", (M, N, K)


def test_unaligned_buffers_fall_back_to_word_access_kernels():
    """the specialised kernels stage operands with bulk copies and store 16-byte vectors; a buffer that is only
    word-aligned (a slice offset by one word) must take the word-access kernels and give the same bits."""
    a = QType(7, 8)
    mul = QType(14, 8, True, RND_CONV, SAT_TCPL)
    lv = [QType(15 + l, 8 - l, True, RND_POS_INF, SAT_TCPL) for l in range(6)]
    c = QType(17, 4, True, RND_POS_INF, SAT_TCPL)
    plan = ops.make_gemul_plan(a, a, c, 64, 64, 64, mul=mul, add=lv)
    batch = 5
    n = batch * 4096
    for offs in ((1, 0, 0), (0, 2, 0), (0, 0, 3), (1, 1, 1)):
        bufs = [torch.zeros(n + 4, dtype=torch.int32, device="cuda") for _ in range(3)]
        A, B, Cd = (b[o:o + n] for b, o in zip(bufs, offs))
        ops.synth_fill(A, a.raw_min, a.raw_max, 7, 0); ops.synth_fill(B, a.raw_min, a.raw_max, 7, 1)
        ops.Qgemul(Cd, A, B, plan, batch=batch)
        want = oracle.gemul_batched(plan, A.cpu().numpy(), B.cpu().numpy(), batch)
        assert np.array_equal(Cd.cpu().numpy(), want), offs
    # complex mat-vec (config-4 types): {re, im} pairs, buffers offset by one complex element (8 bytes)
    t = QType(3, 12); P = QType(6, 12); S = QType(7, 12)
    tags = {"abT": QType(4, 12), "cdT": QType(4, 11), "abcT": P, "cdbT": P, "badT": P, "ABT": S, "BCT": S}
    lvc = [QType(8 + l, 12 - l, True, RND_CONV, SAT_TCPL) for l in range(8)]
    cplan = ops.make_cgemul_plan(t, t, t, t, lvc[-1], lvc[-1], 256, 1, 256, tf_tags=tags, add_re=lvc, form=TREE_VARIADIC)
    nb = 3
    Hb = torch.zeros(nb * 256 * 256 * 2 + 8, dtype=torch.int32, device="cuda"); xb = torch.zeros(nb * 512 + 8, dtype=torch.int32, device="cuda")
    yb = torch.zeros(nb * 512 + 8, dtype=torch.int32, device="cuda")
    for o in (2, 1):                                  # 8-byte aligned only, then 4-byte aligned only
        H, xv, y = Hb[o:o + nb * 256 * 256 * 2], xb[o:o + nb * 512], yb[o:o + nb * 512]
        ops.synth_fill(H, -(1 << 14), (1 << 14) - 1, 3, 0); ops.synth_fill(xv, -(1 << 14), (1 << 14) - 1, 3, 1)
        ops.Qcgemul(y, H, xv, cplan, batch=nb)
        want = oracle.cgemul_batched(cplan, H.cpu().numpy(), xv.cpu().numpy(), nb)
        assert np.array_equal(y.cpu().numpy(), want), o


def test_fp_pipe_quantizers_exhaustive_sweep():
    """the FP32-FMA product quantizer and the FADD pre-adds are a departure from "integer shift/mask/clamp";
    pin them exhaustively.  Config 3: all 2^32 operand pairs through fp_mul_apply (with Qmul's clamp, and clamp-free inside
    the per-instance elision thresholds) against the integer mul_apply.  Config 4: a+b, b-a (2^32 pairs each) and the three
    TFComplexMul products (2^33 + 2^32 + 2^32 pairs) against the integer forms."""
    import ctypes
    lib = abi.load()
    ex = abi.qb_exec(abi.QB_MEM_DEVICE, 0, None, 0, 0)
    a = QType(7, 8)
    mul = QType(14, 8, True, RND_CONV, SAT_TCPL)
    lv = [QType(15 + l, 8 - l, True, RND_POS_INF, SAT_TCPL) for l in range(6)]
    c = QType(17, 4, True, RND_POS_INF, SAT_TCPL)
    plan = ops.make_gemul_plan(a, a, c, 64, 64, 64, mul=mul, add=lv)
    checked, bad = ctypes.c_uint64(), ctypes.c_uint64()
    abi.check(lib.qb200_selftest_fp(0, ctypes.cast(ctypes.byref(plan), ctypes.c_void_p), ctypes.byref(checked), ctypes.byref(bad), ctypes.byref(ex)))
    # 2^32 clamped + nearly 2^32 clamp-free + 2^32 clamp-free on substituted minima (saturation by operand substitution)
    assert checked.value > 2 * (1 << 32) + (1 << 31) and bad.value == 0, (checked.value, bad.value)
    # a TRN::TCPL product (round toward -inf: fma.rm) with another shift
    plan2 = ops.make_gemul_plan(QType(5, 6), QType(7, 4), c, 64, 64, 64, mul=QType(10, 5, True, 5, 0), add=[QType(20, 5)])
    abi.check(lib.qb200_selftest_fp(0, ctypes.cast(ctypes.byref(plan2), ctypes.c_void_p), ctypes.byref(checked), ctypes.byref(bad), ctypes.byref(ex)))
    assert checked.value >= 1 << 24 and bad.value == 0, (checked.value, bad.value)
    # config 4
    t = QType(3, 12); P = QType(6, 12); S = QType(7, 12)
    tags = {"abT": QType(4, 12), "cdT": QType(4, 11), "abcT": P, "cdbT": P, "badT": P, "ABT": S, "BCT": S}
    lvc = [QType(8 + l, 12 - l, True, RND_CONV, SAT_TCPL) for l in range(8)]
    cplan = ops.make_cgemul_plan(t, t, t, t, lvc[-1], lvc[-1], 256, 1, 256, tf_tags=tags, add_re=lvc, form=TREE_VARIADIC)
    abi.check(lib.qb200_selftest_fp(1, ctypes.cast(ctypes.byref(cplan), ctypes.c_void_p), ctypes.byref(checked), ctypes.byref(bad), ctypes.byref(ex)))
    assert checked.value >= 2 * (1 << 32) + (1 << 33) + 2 * (1 << 32) - (1 << 20) and bad.value == 0, (checked.value, bad.value)


@pytest.mark.parametrize("a,b,mul,sform", [
    (QType(3, 4), QType(5, 6), QType(8, 6, True, RND_CONV, SAT_TCPL), False),        # 8 x 12 bits, shift 4
    (QType(7, 8), QType(7, 8), QType(14, 4, True, RND_CONV, SAT_TCPL), True),        # 16 x 16 bits, shift 12, deferred-shift tree
    (QType(9, 2), QType(2, 9), QType(11, 0, True, RND_CONV, SAT_TCPL), False),       # 12 x 12 bits, shift 11 = log2(-min)
    (QType(7, 8), QType(3, 4), QType(14, 2, True, RND_CONV, SAT_TCPL), True)])       # 16 x 8 bits, shift 10 > log2(-min(b)): not substitutable
def test_saturation_by_operand_substitution_other_widths(a, b, mul, sform):
    """qb_device.cuh:fp_subst -- when the only product Qmul's SAT clamp can touch is min x min, the conversion pass replaces
    operands at their type's minimum by a nearby float that makes the plain FMA deliver the saturated value; the register
    tiles then run ONE clamp-free loop.  Other operand widths and shifts than config 3's: the exhaustive GPU sweep of the
    substituted form (third pass of qb200_selftest_fp) and the tile kernel on batches full of minima against the oracle."""
    import ctypes
    lib = abi.load()
    K = 16
    if sform:
        lv = [QType(mul.i + 1, mul.f, True, RND_POS_INF, SAT_TCPL)] + [QType(mul.i + 1 + l, mul.f - l, True, RND_POS_INF, SAT_TCPL) for l in range(1, 4)]
    else:
        lv = [QType(mul.i + 1 + l, mul.f, True, RND_POS_INF, SAT_TCPL) for l in range(4)]
    c = QType(lv[-1].i, lv[-1].f, True, RND_POS_INF, SAT_TCPL)
    plan = ops.make_gemul_plan(a, b, c, 32, 32, K, mul=mul, add=lv)
    info = ops.plan_info_gemul(plan)
    pa, pb, S = (-a.raw_min).bit_length() - 1, (-b.raw_min).bit_length() - 1, a.f + b.f - mul.f
    subst = pa >= S and pb >= S
    ex = abi.qb_exec(abi.QB_MEM_DEVICE, 0, None, 0, 0)
    checked, bad = ctypes.c_uint64(), ctypes.c_uint64()
    abi.check(lib.qb200_selftest_fp(0, ctypes.cast(ctypes.byref(plan), ctypes.c_void_p), ctypes.byref(checked), ctypes.byref(bad), ctypes.byref(ex)))
    pairs = (a.raw_max - a.raw_min + 1) * (b.raw_max - b.raw_min + 1)
    assert bad.value == 0 and (checked.value > 2.5 * pairs) == subst, (checked.value, pairs, bad.value)
    rng = np.random.default_rng(mul.f * 7 + a.i)
    batch = 9
    A = rand_words(rng, a, batch * 32 * K).astype(np.int32); B = rand_words(rng, b, batch * K * 32).astype(np.int32)
    A[rng.random(A.size) < 0.15] = a.raw_min; B[rng.random(B.size) < 0.15] = b.raw_min          # min x min products all over the place
    A[:32 * K] = a.raw_min; B[:K * 32] = b.raw_min                                               # instance 0: nothing but
    Cd = torch.empty(batch * 32 * 32, dtype=torch.int32, device="cuda")
    ops.Qgemul(Cd, dev(A), dev(B), plan, batch=batch)
    assert b"tile" in ops.plan_info_gemul(plan).kernel, info.kernel
    want = oracle.gemul_batched(plan, A, B, batch)
    assert np.array_equal(Cd.cpu().numpy(), want)
    C2 = torch.empty_like(Cd)
    ops.Qgemul(C2, dev(A), dev(B), plan, batch=batch, flags=abi.QB_F_NO_SPECIALIZE)
    assert torch.equal(Cd, C2)


@pytest.mark.parametrize("K,modes,fp", [(64, "PPPPP", True), (64, "PNTPN", True), (32, "TTTT", False), (16, "NPP", True), (8, "PP", False), (64, "PPPPP", False),
                                        (128, "PPPPPP", True), (128, "PNPTPN", True)])
@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
def test_tile_kernel_deferred_shift_tree_vs_oracle(K, modes, fp, ta, tb):
    """The S form of the register-tile kernel (qb_fast_gemul.cuh: exact level 0, clamp-free halving levels with a
    data-independent bias -- RND::POS_INF, RND::NEG_INF, TRN::TCPL at a 1-bit shift -- carried as pre-shift sums with the
    biases folded into compile-time constants) against the oracle and the runtime-descriptor kernel, full-range operands,
    instances with and without the saturating min x min product."""
    from qublas_b200.qtypes import RND_NEG_INF, TRN_TCPL
    m = {"P": RND_POS_INF, "N": RND_NEG_INF, "T": TRN_TCPL}
    a = QType(7, 8)
    mul = QType(14, 8, True, RND_CONV if fp else RND_POS_INF, SAT_TCPL)          # CONV: FP32-FMA product; POS_INF: integer product
    lv = [QType(15, 8, True, RND_POS_INF, SAT_TCPL)] + [QType(15 + l, 8 - l, True, m[modes[l - 1]], SAT_TCPL) for l in range(1, len(modes) + 1)]
    assert len(lv) == K.bit_length() - 1
    c = QType(17, 4, True, RND_POS_INF, SAT_TCPL)
    if (ta or tb) and not (fp and K in (64, 128)):
        pytest.skip("transposed operands: the substituted / negated forms only")
    plan = ops.make_gemul_plan(a, a, c, 64, 32, K, mul=mul, add=lv, trans_a=ta, trans_b=tb)
    batch = 11
    A = torch.empty(batch * 64 * K, dtype=torch.int32, device="cuda"); B = torch.empty(batch * K * 32, dtype=torch.int32, device="cuda")
    ops.synth_fill(A, a.raw_min, a.raw_max, 21, 0); ops.synth_fill(B, a.raw_min, a.raw_max, 21, 1)
    A[:64 * K] = a.raw_min; B[:K * 32] = a.raw_min                                # instance 0: every product saturates
    A[3 * 64 * K + 5] = a.raw_min; B[3 * K * 32 + 7] = a.raw_min                  # instance 3: the clamped variant on random data
    Cd = torch.empty(batch * 64 * 32, dtype=torch.int32, device="cuda")
    ops.Qgemul(Cd, A, B, plan, batch=batch)
    assert b"tile" in ops.plan_info_gemul(plan).kernel
    want = oracle.gemul_batched(plan, A.cpu().numpy(), B.cpu().numpy(), batch)
    assert np.array_equal(Cd.cpu().numpy(), want)
    C2 = torch.empty_like(Cd)
    ops.Qgemul(C2, A, B, plan, batch=batch, flags=abi.QB_F_NO_SPECIALIZE)
    assert torch.equal(Cd, C2)


@pytest.mark.parametrize("M,N,K", [(64, 64, 64), (128, 64, 32), (64, 32, 128), (32, 64, 16)])
@pytest.mark.parametrize("product", ["tf", "basic", "basic_q"])
def test_complex_gemul_register_tile_kernel_vs_oracle(M, N, K, product):
    """Batched complex Qgemul with N > 1 on the register-tile kernel (qb_fast_cgemul_tile.cuh): TFComplexMul and
    BasicComplexMul products with distinct node types, full-range operands, a shared B operand; same bits as the oracle and as
    the runtime-descriptor kernel."""
    t = QType(3, 12); P = QType(6, 12); S = QType(7, 12)
    tags = {"abT": QType(4, 12), "cdT": QType(4, 11), "abcT": P, "cdbT": P, "badT": P, "ABT": S, "BCT": S} if product == "tf" else None
    if product == "basic_q":           # quantized BasicComplexMul products (FP pipe); "basic" keeps the full-precision defaults (integer products)
        tags = {"acT": P, "bdT": QType(6, 11, True, RND_CONV, SAT_TCPL), "adT": P, "bcT": QType(5, 12), "acbdT": S, "adbcT": S}
        product = "basic"
    lv = [QType(8 + l, 12 - l, True, RND_CONV, SAT_TCPL) for l in range((K - 1).bit_length())]
    lvi = [QType(9 + l, 11 - l, True, RND_POS_INF, SAT_TCPL) for l in range((K - 1).bit_length())]
    plan = ops.make_cgemul_plan(t, t, t, t, lv[-1], QType(12, 3), M, N, K, tf_tags=tags, add_re=lv, add_im=lvi, form=TREE_VARIADIC, product=product)
    batch = 7
    A = torch.empty(batch * M * K * 2, dtype=torch.int32, device="cuda"); B = torch.empty(batch * K * N * 2, dtype=torch.int32, device="cuda")
    ops.synth_fill(A, t.raw_min, t.raw_max, 31, 0); ops.synth_fill(B, t.raw_min, t.raw_max, 31, 1)
    A[:64] = t.raw_min; B[:64] = t.raw_min
    Cd = torch.empty(batch * M * N * 2, dtype=torch.int32, device="cuda")
    ops.Qcgemul(Cd, A, B, plan, batch=batch)
    assert b"tile" in ops.plan_info_cgemul(plan).kernel
    want = oracle.cgemul_batched(plan, A.cpu().numpy(), B.cpu().numpy(), batch)
    assert np.array_equal(Cd.cpu().numpy(), want)
    C2 = torch.empty_like(Cd)
    ops.Qcgemul(C2, A, B, plan, batch=batch, flags=abi.QB_F_NO_SPECIALIZE)
    assert torch.equal(Cd, C2)
    plan.stride_b = 0                                      # one B for all instances
    ops.Qcgemul(Cd, A, B, plan, batch=batch)
    want = oracle.cgemul_batched(plan, A.cpu().numpy(), B.cpu().numpy(), batch)
    assert np.array_equal(Cd.cpu().numpy(), want)
