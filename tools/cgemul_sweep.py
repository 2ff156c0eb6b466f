"""Batched complex Qgemul with N > 1 (BASELINE config-4 types): which kernel the library picks and what it delivers, checked
against the oracle on a few instances.  QB200_NO_CTILE=1 selects the thread-per-output kernel for comparison.
python tools/cgemul_sweep.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle  # noqa: E402
from qublas_b200 import ops  # noqa: E402
from qublas_b200.qtypes import QType, RND_CONV, SAT_TCPL, TREE_VARIADIC  # noqa: E402


def main():
    t = QType(3, 12); P = QType(6, 12); S = QType(7, 12)
    tags = {"abT": QType(4, 12), "cdT": QType(4, 11), "abcT": P, "cdbT": P, "badT": P, "ABT": S, "BCT": S}
    for (M, N, K) in ((64, 64, 64), (32, 32, 32), (64, 16, 128), (128, 64, 32), (16, 16, 16)):
        for prod in ("tf", "basic"):
            lv = [QType(8 + l, 12 - l, True, RND_CONV, SAT_TCPL) for l in range((K - 1).bit_length())]
            btags = {"acT": P, "bdT": P, "adT": P, "bcT": P, "acbdT": S, "adbcT": S}       # quantized BasicComplexMul products (the defaults keep full precision)
            plan = ops.make_cgemul_plan(t, t, t, t, lv[-1], lv[-1], M, N, K, tf_tags=tags if prod == "tf" else btags, add_re=lv, form=TREE_VARIADIC, product=prod)
            words = 2 * (M * K + K * N + M * N)
            batch = max(64, min(1 << 18, (1 << 27) // words))
            A = torch.empty(batch * M * K * 2, dtype=torch.int32, device="cuda"); B = torch.empty(batch * K * N * 2, dtype=torch.int32, device="cuda")
            ops.synth_fill(A, -(1 << 14), (1 << 14) - 1, 1, 0); ops.synth_fill(B, -(1 << 14), (1 << 14) - 1, 1, 1)
            Cd = torch.empty(batch * M * N * 2, dtype=torch.int32, device="cuda")
            for _ in range(2):
                ops.Qcgemul(Cd, A, B, plan, batch=batch)
            torch.cuda.synchronize()
            n = 3
            want = oracle.cgemul_batched(plan, A[:n * M * K * 2].cpu().numpy(), B[:n * K * N * 2].cpu().numpy(), n)
            ok = bool(np.array_equal(Cd[:n * M * N * 2].cpu().numpy(), want))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            e0.record()
            for _ in range(reps):
                ops.Qcgemul(Cd, A, B, plan, batch=batch)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            print(json.dumps({"shape": [M, N, K], "product": prod, "batch": batch, "kernel": ops.plan_info_cgemul(plan).kernel.decode(), "bit_exact": ok,
                              "g_cmacs": round(batch * M * N * K / ms / 1e6, 1), "gbs": round(words * 4 * batch / ms / 1e6, 1)}), flush=True)
            del A, B, Cd


if __name__ == "__main__":
    main()
