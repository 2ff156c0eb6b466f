#!/usr/bin/env python3
"""BASELINE config 5 on every visible GPU from ONE process (qb200_gemul_sharded: M-row panels, B replicated, no traffic on the
compute path) and, as a separate line, the optional all-gather of the C panels over NVLink / NVSwitch (qb200_allgather:
2-D peer copies re-pitching each panel into the column-major C of every device).  SURVEY.md 8(e): the gather is reported
on its own, never folded into the throughput.   python tools/gather_probe.py [M=N=K] [int8]"""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qublas_b200 import abi, ops
from qublas_b200.qtypes import QType

Mt = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
int8 = len(sys.argv) > 2 and sys.argv[2] == "int8"
n = torch.cuda.device_count()
lib = abi.load()
a = QType(7, 0); acc = QType(30, 0)
rows = Mt // n
assert rows * n == Mt and rows % 256 == 0, "M must split into panels of a multiple of 256 rows"
wbytes = (1, 1, 0) if int8 else (0, 0, 0)
plan = ops.make_gemul_plan(a, a, acc, rows, Mt, Mt, mul=QType(15, 0), add=[acc], trans_a=True, word_bytes=wbytes)   # A stored K-major: a row panel is one contiguous block
dt = torch.int8 if int8 else torch.int32
A, B, P, Cg = [], [], [], []
for d in range(n):
    dev = torch.device("cuda", d)
    Ad = torch.empty(rows * Mt, dtype=torch.int32, device=dev); Bd = torch.empty(Mt * Mt, dtype=torch.int32, device=dev)
    ops.synth_fill(Ad, -128, 127, 1, 0, d * rows * Mt); ops.synth_fill(Bd, -128, 127, 1, 1)
    A.append(Ad.to(dt)); B.append(Bd.to(dt))
    P.append(torch.empty(rows * Mt, dtype=torch.int32, device=dev))            # the panel's rows of C, dense (leading dimension = rows)
    Cg.append(torch.empty(Mt * Mt, dtype=torch.int32, device=dev))             # the whole C on every device (only for the gather)
shards = (abi.qb_shard * n)(*[abi.qb_shard(d, 0, A[d].data_ptr(), B[d].data_ptr(), P[d].data_ptr(), 1) for d in range(n)])
ex = abi.qb_exec(abi.QB_MEM_DEVICE, -1, None, 0, 0)
vp, u64 = C.c_void_p, C.c_uint64
dst = (vp * n)(*[t.data_ptr() for t in Cg]); src = (vp * n)(*[t.data_ptr() for t in P]); dv = (C.c_int32 * n)(*range(n))
off = (u64 * n)(*[d * rows * 4 for d in range(n)]); sp = (u64 * n)(*[rows * 4] * n); wid = (u64 * n)(*[rows * 4] * n)


def sync():
    for d in range(n):
        torch.cuda.synchronize(d)


for _ in range(3):
    abi.check(lib.qb200_gemul_sharded(C.byref(plan), shards, n, C.byref(ex)))
sync()
reps = 20
t0 = time.perf_counter()
for _ in range(reps):
    abi.check(lib.qb200_gemul_sharded(C.byref(plan), shards, n, C.byref(ex)))
sync()
gemm_s = (time.perf_counter() - t0) / reps
sec = C.c_double(); best = 1e9
for _ in range(5):
    abi.check(lib.qb200_allgather(dst, src, dv, n, Mt, Mt * 4, off, sp, wid, C.byref(sec)))
    best = min(best, sec.value)
# the same in ONE kernel per GPU: the epilogue stores into every device's C (qb200_gemul_panels_gathered)
cf = (vp * n)(*[t.data_ptr() for t in Cg])
for t in Cg:
    t.zero_()
sync()
for _ in range(3):
    abi.check(lib.qb200_gemul_panels_gathered(C.byref(plan), shards, n, cf, C.byref(ex)))
sync()
t0 = time.perf_counter()
for _ in range(reps):
    abi.check(lib.qb200_gemul_panels_gathered(C.byref(plan), shards, n, cf, C.byref(ex)))
sync()
fused_s = (time.perf_counter() - t0) / reps
# check: device 0's gathered C against an exact float64 GEMM on sampled rows of every panel
ok = True
for d in range(n):
    r = torch.arange(0, rows, 257, device="cuda:0")
    Ad = A[d].to("cuda:0").view(rows, Mt)[r].double(); Bd = B[0].view(Mt, Mt).t().double()
    want = Ad @ Bd
    ok = ok and bool(torch.equal(Cg[0].view(Mt, Mt).t()[d * rows + r].double(), want)) and bool(torch.equal(Cg[n - 1].view(Mt, Mt).t()[d * rows + r.to(Cg[n - 1].device)].double().to("cuda:0"), want))
ops_total = 2.0 * Mt * Mt * Mt
print(json.dumps({"workload": "config 5, %d^3, M-row panels of %d rows on %d GPU(s), %s operands" % (Mt, rows, n, "int8 K-major (in place)" if int8 else "int32 words"),
                  "gemm_ms_wall_per_call": gemm_s * 1e3, "gemm_tops_total": ops_total / gemm_s / 1e12,
                  "allgather_ms_device": best * 1e3, "allgather_bytes_per_gpu_received": (n - 1) * rows * Mt * 4,
                  "allgather_gbs_per_gpu": (n - 1) * rows * Mt * 4 / best / 1e9 if n > 1 else None,
                  "gemm_plus_allgather_ms": gemm_s * 1e3 + best * 1e3, "fused_gemm_allgather_ms_wall_per_call": fused_s * 1e3,
                  "gathered_c_checked": ok}))
