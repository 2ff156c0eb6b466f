import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from oracle import oracle
from qublas_b200 import ops
from qublas_b200.qtypes import QType, RND_CONV, RND_POS_INF, SAT_TCPL
a = QType(7, 8); mul = QType(14, 8, True, RND_CONV, SAT_TCPL); c = QType(17, 4, True, RND_POS_INF, SAT_TCPL)
for (M, N, K) in [(64, 64, 128), (32, 32, 128), (64, 64, 256)]:
    nl = (K - 1).bit_length()
    lv = [QType(15 + l, max(8 - l, 1), True, RND_POS_INF, SAT_TCPL) for l in range(nl)]
    plan = ops.make_gemul_plan(a, a, c, M, N, K, mul=mul, add=lv)
    batch = (1 << 28) // (M * K + K * N + M * N)
    A = torch.empty(batch * M * K, dtype=torch.int32, device="cuda"); B = torch.empty(batch * K * N, dtype=torch.int32, device="cuda")
    ops.synth_fill(A, a.raw_min, a.raw_max, 1, 0); ops.synth_fill(B, a.raw_min, a.raw_max, 1, 1)
    A[:50] = a.raw_min; B[:50] = a.raw_min
    Cd = torch.empty(batch * M * N, dtype=torch.int32, device="cuda")
    ops.Qgemul(Cd, A, B, plan, batch=batch)
    n = 6
    want = oracle.gemul_batched(plan, A[:n * M * K].cpu().numpy(), B[:n * K * N].cpu().numpy(), n)
    ok = bool(np.array_equal(Cd[:n * M * N].cpu().numpy(), want))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ops.Qgemul(Cd, A, B, plan, batch=batch)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(json.dumps({"shape": [M, N, K], "defines": os.environ.get("QB200_JIT_DEFINES", ""), "kernel": ops.plan_info_gemul(plan).kernel.decode(), "bit_exact": ok, "gmacs": round(batch * M * N * K / ms / 1e6, 1)}))
