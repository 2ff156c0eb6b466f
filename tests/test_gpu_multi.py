"""One call, several GPUs: exec.device = QB_DEVICE_ALL on host buffers, device-resident shards through
qb200_gemul_sharded, and the optional output all-gather (qb200_allgather: contiguous batch shards and M-row panels of a
column-major matrix).  Every test also runs on a 1-GPU box (the shards then share device 0, the peer copies become
device-to-device copies); with >= 2 GPUs the same code spans them."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle
from qublas_b200 import abi, ops
from qublas_b200.qtypes import QType, RND_CONV, RND_POS_INF, SAT_TCPL

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def c3_plan(M=64, N=64, K=64):
    a = QType(7, 8)
    mul = QType(14, 8, True, RND_CONV, SAT_TCPL)
    lv = [QType(15 + l, 8 - l, True, RND_POS_INF, SAT_TCPL) for l in range(6)]
    c = QType(17, 4, True, RND_POS_INF, SAT_TCPL)
    return a, ops.make_gemul_plan(a, a, c, M, N, K, mul=mul, add=lv)


def test_device_all_shards_host_buffers_over_every_gpu():
    a, plan = c3_plan()
    batch = 203                                   # not a multiple of any GPU count
    A = oracle.synth_fill(batch * 4096, 4, a.raw_min, a.raw_max, 9, 0); B = oracle.synth_fill(batch * 4096, 4, a.raw_min, a.raw_max, 9, 1)
    Cc = np.zeros(batch * 4096, dtype=np.int32)
    ex = abi.qb_exec(abi.QB_MEM_HOST, abi.QB_DEVICE_ALL, None, 0, 0)
    lib = abi.load()
    abi.check(lib.qb200_gemul_batched(C.byref(plan), A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p), Cc.ctypes.data_as(C.c_void_p), batch, C.byref(ex)))
    assert np.array_equal(Cc, oracle.gemul_batched(plan, A, B, batch))
    # Qreduce: 37 vectors of 5000 elements
    elem = QType(8, 8); lv = [QType(9 + l, 8 - l, True, RND_POS_INF, SAT_TCPL) for l in range(8)]
    rp = ops.make_reduce_plan(elem, lv, 5000)
    x = oracle.synth_fill(37 * 5000, 4, elem.raw_min, elem.raw_max, 9, 2)
    out = np.zeros(37, dtype=np.int32)
    abi.check(lib.qb200_reduce_batched(C.byref(rp), x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), None, 0, 37, C.byref(ex)))
    want, _ = oracle.reduce_batched(rp, x, 37)
    assert np.array_equal(out, want)
    # device memory cannot be sharded implicitly
    exd = abi.qb_exec(abi.QB_MEM_DEVICE, abi.QB_DEVICE_ALL, None, 0, 0)
    assert lib.qb200_gemul_batched(C.byref(plan), A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p), Cc.ctypes.data_as(C.c_void_p), batch, C.byref(exd)) == abi.QB_EINVAL


def _devices(n):
    have = torch.cuda.device_count()
    return [i % have for i in range(n)]


@pytest.mark.parametrize("n_shards", [2, 3])
def test_sharded_device_resident_batches_and_contiguous_allgather(n_shards):
    a, plan = c3_plan()
    lib = abi.load()
    devs = _devices(n_shards)
    total = 50
    first = [0] * n_shards; count = [0] * n_shards
    for r in range(n_shards):
        f, c = C.c_uint64(), C.c_uint64()
        abi.check(lib.qb200_shard_range(total, n_shards, r, C.byref(f), C.byref(c)))
        first[r], count[r] = f.value, c.value
    A, B, Cfull = [], [], []
    shards = (abi.qb_shard * n_shards)()
    for r, d in enumerate(devs):
        dev = torch.device("cuda", d)
        Ar = torch.empty(count[r] * 4096, dtype=torch.int32, device=dev); Br = torch.empty_like(Ar)
        ops.synth_fill(Ar, a.raw_min, a.raw_max, 4, 0, first[r] * 4096); ops.synth_fill(Br, a.raw_min, a.raw_max, 4, 1, first[r] * 4096)
        Cr = torch.zeros(total * 4096, dtype=torch.int32, device=dev)           # full-size output buffer on every device
        A.append(Ar); B.append(Br); Cfull.append(Cr)
        shards[r] = abi.qb_shard(d, 0, Ar.data_ptr(), Br.data_ptr(), Cr.data_ptr() + first[r] * 4096 * 4, count[r])
    torch.cuda.synchronize()
    ex = abi.qb_exec(abi.QB_MEM_DEVICE, -1, None, 0, 0)
    abi.check(lib.qb200_gemul_sharded(C.byref(plan), shards, n_shards, C.byref(ex)))
    hA = oracle.synth_fill(total * 4096, 4, a.raw_min, a.raw_max, 4, 0); hB = oracle.synth_fill(total * 4096, 4, a.raw_min, a.raw_max, 4, 1)
    want = oracle.gemul_batched(plan, hA, hB, total)
    for r in range(n_shards):
        sl = slice(first[r] * 4096, (first[r] + count[r]) * 4096)
        assert np.array_equal(Cfull[r][sl].cpu().numpy(), want[sl])
    # optional all-gather: every device ends up with all of C
    vp, u64 = C.c_void_p, C.c_uint64
    dst = (vp * n_shards)(*[t.data_ptr() for t in Cfull])
    src = (vp * n_shards)(*[Cfull[r].data_ptr() + first[r] * 4096 * 4 for r in range(n_shards)])
    dv = (C.c_int32 * n_shards)(*devs)
    off = (u64 * n_shards)(*[first[r] * 4096 * 4 for r in range(n_shards)])
    wid = (u64 * n_shards)(*[count[r] * 4096 * 4 for r in range(n_shards)])
    sec = C.c_double()
    abi.check(lib.qb200_allgather(dst, src, dv, n_shards, 1, 0, off, None, wid, C.byref(sec)))
    assert sec.value > 0
    for r in range(n_shards):
        assert np.array_equal(Cfull[r].cpu().numpy(), want), r


def test_allgather_of_row_panels_into_column_major_matrices():
    """M-row panels (config 5's sharding): panel r = rows [r0, r1) of a column-major M x N int32 matrix, computed dense on
    its device, lands in the global column-major C of every device (2-D peer copies re-pitch it)."""
    lib = abi.load()
    n = 3
    devs = _devices(n)
    M, N = 384, 128
    rows = [128, 192, 64]; r0 = [0, 128, 320]
    rng = np.random.default_rng(3)
    full = rng.integers(-2 ** 30, 2 ** 30, size=(N, M), dtype=np.int32)          # [column][row] = column-major M x N
    panels = [torch.from_numpy(np.ascontiguousarray(full[:, r0[r]:r0[r] + rows[r]])).to(torch.device("cuda", devs[r])) for r in range(n)]
    Cg = [torch.full((N, M), -1, dtype=torch.int32, device=torch.device("cuda", devs[r])) for r in range(n)]
    vp, u64 = C.c_void_p, C.c_uint64
    dst = (vp * n)(*[t.data_ptr() for t in Cg]); src = (vp * n)(*[p.data_ptr() for p in panels])
    dv = (C.c_int32 * n)(*devs)
    off = (u64 * n)(*[r0[r] * 4 for r in range(n)]); sp = (u64 * n)(*[rows[r] * 4 for r in range(n)]); wid = (u64 * n)(*[rows[r] * 4 for r in range(n)])
    torch.cuda.synchronize()
    abi.check(lib.qb200_allgather(dst, src, dv, n, N, M * 4, off, sp, wid, None))
    for r in range(n):
        assert np.array_equal(Cg[r].cpu().numpy(), full), r


@pytest.mark.parametrize("n,words", [(2, 4), (3, 4), (2, 1)])
def test_gemm_and_allgather_in_one_kernel_per_gpu(n, words):
    """qb200_gemul_panels_gathered: panel r of a (n * M) x N Qgemul on the tensor-core path, whose epilogue stores every
    element into the column-major C of EVERY device (local + peer stores) -- no copy pass.  Every device's C equals the exact
    float64 GEMM of the whole problem; int32 words (narrowed inside the same kernel) and int8 operands read in place."""
    lib = abi.load()
    devs = _devices(n)
    M, N, K = 256, 512, 384
    t = QType(7, 0); acc = QType(30, 0)
    plan = ops.make_gemul_plan(t, t, acc, M, N, K, mul=QType(15, 0), add=[acc], trans_a=True, word_bytes=(words if words == 1 else 0, words if words == 1 else 0, 0))
    assert ops.plan_info_gemul(plan).tensor_core == 1
    rng = np.random.default_rng(n + words)
    dt = np.int8 if words == 1 else np.int32
    Afull = rng.integers(-128, 128, size=(n * M, K)).astype(dt)                 # op(A) row-major = stored K-major
    Bm = rng.integers(-128, 128, size=(N, K)).astype(dt)                        # B stored K x N column-major = [j][k]
    want = (Afull.astype(np.float64) @ Bm.astype(np.float64).T).T               # [j][i] = column-major (n * M) x N
    A, B, Cf = [], [], []
    for r, d in enumerate(devs):
        dev = torch.device("cuda", d)
        A.append(torch.from_numpy(np.ascontiguousarray(Afull[r * M:(r + 1) * M])).to(dev)); B.append(torch.from_numpy(Bm).to(dev))
        Cf.append(torch.full((N, n * M), -7, dtype=torch.int32, device=dev))
    shards = (abi.qb_shard * n)(*[abi.qb_shard(devs[r], 0, A[r].data_ptr(), B[r].data_ptr(), 0, 1) for r in range(n)])
    cf = (C.c_void_p * n)(*[x.data_ptr() for x in Cf])
    for d in set(devs):
        torch.cuda.synchronize(d)
    ex = abi.qb_exec(abi.QB_MEM_DEVICE, -1, None, 0, 0)
    for _ in range(2):                                                          # twice: flags of the fused narrowing are per call
        abi.check(lib.qb200_gemul_panels_gathered(C.byref(plan), shards, n, cf, C.byref(ex)))
    for r in range(n):
        assert np.array_equal(Cf[r].cpu().numpy().astype(np.float64), want), r
    # plans off the tensor-core path are refused
    a3, p3 = c3_plan()
    assert lib.qb200_gemul_panels_gathered(C.byref(p3), shards, n, cf, C.byref(ex)) == abi.QB_EINVAL
