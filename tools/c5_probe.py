#!/usr/bin/env python3
"""Config-5 call timed three ways: isolated calls (the GPU idles between them: boost clocks, no power cap), a sustained
back-to-back loop (what bench.py times), for int32 words (narrowed inside the GEMM kernel, or by separate pack launches with
QB200_I8_NO_FUSE=1) and for int8 K-major operands read in place; cuBLASLt int8 (torch._int_mm) beside them.
usage: python tools/c5_probe.py [M]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qublas_b200 import ops
from qublas_b200.qtypes import QType

M = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda", 0)
a = QType(7, 0); acc = QType(30, 0)
plan32 = ops.make_gemul_plan(a, a, acc, M, M, M, mul=QType(15, 0), add=[acc], trans_a=True)
plan8 = ops.make_gemul_plan(a, a, acc, M, M, M, mul=QType(15, 0), add=[acc], trans_a=True, word_bytes=(1, 1, 0))
A = torch.empty(M * M, dtype=torch.int32, device=dev); B = torch.empty_like(A); Cd = torch.empty_like(A)
ops.synth_fill(A, -128, 127, 1, 0); ops.synth_fill(B, -128, 127, 1, 1)
A8, B8 = A.to(torch.int8), B.to(torch.int8)
A8m, B8m = A8.view(M, M), B8.view(M, M).t()
ops_total = 2.0 * M * M * M


def isolated(fn, n=12):
    ts = []
    for _ in range(n):
        time.sleep(0.05)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def sustained(fn, n=60):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


cases = {"int32 words": lambda: ops.Qgemul(Cd, A, B, plan32), "int8 in place": lambda: ops.Qgemul(Cd, A8, B8, plan8),
         "cuBLASLt int8": lambda: torch._int_mm(A8m, B8m)}
only = os.environ.get("C5_ONLY")
for name, fn in cases.items():
    if only and only not in name:
        continue
    fn(); torch.cuda.synchronize()
    lo, med = isolated(fn)
    sus = sustained(fn)
    print(json.dumps({"case": name, "fuse": os.environ.get("QB200_I8_NO_FUSE", "0") != "1", "pk": os.environ.get("QB200_I8_PK", ""), "M": M, "isolated_min_ms": round(lo, 4), "isolated_median_ms": round(med, 4),
                      "sustained_ms": round(sus, 4), "isolated_pops": round(ops_total / lo / 1e12, 1), "sustained_pops": round(ops_total / sus / 1e12, 1)}), flush=True)
