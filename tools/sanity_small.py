#!/usr/bin/env python3
"""Tiny run of every kernel family (for compute-sanitizer): small shapes, checked against the oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import oracle
from qublas_b200 import abi, ops
from qublas_b200.qtypes import QType, RND_CONV, RND_POS_INF, SAT_TCPL, TREE_VARIADIC

dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
a = QType(7, 8); mul = QType(14, 8, True, RND_CONV, SAT_TCPL)
lv = [QType(15 + l, 8 - l, True, RND_POS_INF, SAT_TCPL) for l in range(6)]
c = QType(17, 4, True, RND_POS_INF, SAT_TCPL)
# specialised tile kernel (TMA ring) + generic kernel
plan = ops.make_gemul_plan(a, a, c, 64, 64, 64, mul=mul, add=lv)
A = torch.empty(5 * 4096, dtype=torch.int32, device=dev); B = torch.empty_like(A); C = torch.empty_like(A)
ops.synth_fill(A, a.raw_min, a.raw_max, 1, 0); ops.synth_fill(B, a.raw_min, a.raw_max, 1, 1)
ops.Qgemul(C, A, B, plan, batch=5)
assert np.array_equal(C.cpu().numpy(), oracle.gemul_batched(plan, A.cpu().numpy(), B.cpu().numpy(), 5))
ops.Qgemul(C, A, B, plan, batch=5, flags=abi.QB_F_NO_SPECIALIZE)
# JIT tile kernel
plan2 = ops.make_gemul_plan(a, a, c, 32, 32, 16, mul=QType(14, 8, True, 3, 2), add=lv[:4])
A2 = A[:3 * 512]; B2 = B[:3 * 512]; C2 = torch.empty(3 * 1024, dtype=torch.int32, device=dev)
ops.Qgemul(C2, A2, B2, plan2, batch=3)
assert np.array_equal(C2.cpu().numpy(), oracle.gemul_batched(plan2, A2.cpu().numpy(), B2.cpu().numpy(), 3))
# reduce: specialised-size JIT (2^16), generic odd length, thread kernel
elem = QType(8, 8); lr = [QType(9 + l, 8 - l, True, l % 7, SAT_TCPL) for l in range(8)]
x = torch.empty(2 * 70001, dtype=torch.int32, device=dev); ops.synth_fill(x, elem.raw_min, elem.raw_max, 1, 2)
for n in (1 << 16, 70001, 37):
    out, _ = ops.Qreduce(x[:2 * n], elem, lr, length=n, batch=2)
    want, _ = oracle.reduce_batched(ops.make_reduce_plan(elem, lr, n), x[:2 * n].cpu().numpy(), 2)
    assert np.array_equal(out.cpu().numpy(), want), n
# complex mat-vec (AOT C4 kernel) and generic complex
t = QType(3, 12); P = QType(6, 12); S = QType(7, 12)
tags = {"abT": QType(4, 12), "cdT": QType(4, 11), "abcT": P, "cdbT": P, "badT": P, "ABT": S, "BCT": S}
lc = [QType(8 + l, 12 - l, True, RND_CONV, SAT_TCPL) for l in range(8)]
pc = ops.make_cgemul_plan(t, t, t, t, lc[-1], lc[-1], 256, 1, 256, tf_tags=tags, add_re=lc, form=TREE_VARIADIC)
H = torch.empty(2 * 65536 * 2, dtype=torch.int32, device=dev); xv = torch.empty(2 * 512, dtype=torch.int32, device=dev)
ops.synth_fill(H, t.raw_min, t.raw_max, 1, 3); ops.synth_fill(xv, t.raw_min, t.raw_max, 1, 4)
y = torch.empty(2 * 512, dtype=torch.int32, device=dev)
ops.Qcgemul(y, H, xv, pc, batch=2)
assert np.array_equal(y.cpu().numpy(), oracle.cgemul_batched(pc, H.cpu().numpy(), xv.cpu().numpy(), 2))
# tcgen05 path (both variants by shape parity)
for M in (128, 256):
    a8 = QType(7, 0); acc = QType(30, 0)
    p8 = ops.make_gemul_plan(a8, a8, acc, M, 256, 256, mul=QType(15, 0), add=[acc])
    A8 = torch.from_numpy(rng.integers(-128, 128, M * 256).astype(np.int32)).to(dev); B8 = torch.from_numpy(rng.integers(-128, 128, 256 * 256).astype(np.int32)).to(dev)
    C8 = torch.empty(M * 256, dtype=torch.int32, device=dev)
    ops.Qgemul(C8, A8, B8, p8)
    assert np.array_equal(C8.cpu().numpy(), oracle.gemul_batched(p8, A8.cpu().numpy(), B8.cpu().numpy(), 1))
torch.cuda.synchronize()
print("sanity ok")
