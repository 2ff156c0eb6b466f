#!/usr/bin/env python3
"""Small driver for ncu captures: runs one workload's hot kernel a few times on device-resident data.
usage: python tools/prof.py c2|c3|c4|c5|c5t|c5i8|cg [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qublas_b200 import ops
from qublas_b200.qtypes import QType, RND_CONV, RND_POS_INF, SAT_TCPL, TREE_VARIADIC

w = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
if w == "c2":
    n, batch = 1 << 24, 16
    elem = QType(8, 8)
    lv = [QType(9 + l, 8 - l, True, RND_POS_INF, SAT_TCPL) for l in range(8)]
    lv += [QType(17 + (l - 7 + 1) // 2, 0, True, RND_POS_INF, SAT_TCPL) for l in range(8, 24)]
    x = torch.empty(batch * n, dtype=torch.int32, device=dev)
    ops.synth_fill(x, elem.raw_min, elem.raw_max, 1, 0)
    fn = lambda: ops.Qreduce(x, elem, lv, length=n, batch=batch)
elif w == "c3":
    a = QType(7, 8); mul = QType(14, 8, True, RND_CONV, SAT_TCPL)
    lv = [QType(15 + l, 8 - l, True, RND_POS_INF, SAT_TCPL) for l in range(6)]
    c = QType(17, 4, True, RND_POS_INF, SAT_TCPL)
    plan = ops.make_gemul_plan(a, a, c, 64, 64, 64, mul=mul, add=lv)
    batch = 32768
    A = torch.empty(batch * 4096, dtype=torch.int32, device=dev); B = torch.empty_like(A); Cd = torch.empty_like(A)
    ops.synth_fill(A, a.raw_min, a.raw_max, 1, 0); ops.synth_fill(B, a.raw_min, a.raw_max, 1, 1)
    fn = lambda: ops.Qgemul(Cd, A, B, plan, batch=batch)
elif w == "c4":
    t = QType(3, 12); P = QType(6, 12); S = QType(7, 12)
    tags = {"abT": QType(4, 12), "cdT": QType(4, 11), "abcT": P, "cdbT": P, "badT": P, "ABT": S, "BCT": S}
    lv = [QType(8 + l, 12 - l, True, RND_CONV, SAT_TCPL) for l in range(8)]
    plan = ops.make_cgemul_plan(t, t, t, t, lv[-1], lv[-1], 256, 1, 256, tf_tags=tags, add_re=lv, form=TREE_VARIADIC)
    batch = 4096
    H = torch.empty(batch * 65536 * 2, dtype=torch.int32, device=dev); xv = torch.empty(batch * 512, dtype=torch.int32, device=dev)
    ops.synth_fill(H, -(1 << 14), (1 << 14) - 1, 1, 0); ops.synth_fill(xv, -(1 << 14), (1 << 14) - 1, 1, 1)
    y = torch.empty(batch * 512, dtype=torch.int32, device=dev)
    fn = lambda: ops.Qcgemul(y, H, xv, plan, batch=batch)
elif w == "cg":                                   # batched complex 64x64x64 TFComplexMul GEMM (register-tile kernel, FP products)
    t = QType(3, 12); P = QType(6, 12); S = QType(7, 12)
    tags = {"abT": QType(4, 12), "cdT": QType(4, 11), "abcT": P, "cdbT": P, "badT": P, "ABT": S, "BCT": S}
    lv = [QType(8 + l, 12 - l, True, RND_CONV, SAT_TCPL) for l in range(6)]
    plan = ops.make_cgemul_plan(t, t, t, t, lv[-1], lv[-1], 64, 64, 64, tf_tags=tags, add_re=lv, form=TREE_VARIADIC)
    batch = 4096
    A = torch.empty(batch * 8192, dtype=torch.int32, device=dev); B = torch.empty_like(A); Cd = torch.empty_like(A)
    ops.synth_fill(A, -(1 << 14), (1 << 14) - 1, 1, 0); ops.synth_fill(B, -(1 << 14), (1 << 14) - 1, 1, 1)
    fn = lambda: ops.Qcgemul(Cd, A, B, plan, batch=batch)
else:
    M = 8192
    a = QType(7, 0); acc = QType(30, 0)
    plan = ops.make_gemul_plan(a, a, acc, M, M, M, mul=QType(15, 0), add=[acc], trans_a=(w == "c5t"))     # c5t: A stored K x M as in bench.py
    A = torch.empty(M * M, dtype=torch.int32, device=dev); B = torch.empty_like(A); Cd = torch.empty_like(A)
    ops.synth_fill(A, -128, 127, 1, 0); ops.synth_fill(B, -128, 127, 1, 1)
    fn = lambda: ops.Qgemul(Cd, A, B, plan)
    if w == "c5i8":                               # int8 words, K contiguous: read in place by TMA (no narrowing at all)
        plan = ops.make_gemul_plan(a, a, acc, M, M, M, mul=QType(15, 0), add=[acc], trans_a=True, word_bytes=(1, 1, 0))
        A8, B8 = A.to(torch.int8), B.to(torch.int8)
        fn = lambda: ops.Qgemul(Cd, A8, B8, plan)
for _ in range(iters):
    fn()
torch.cuda.synchronize()
print("ok", w)
