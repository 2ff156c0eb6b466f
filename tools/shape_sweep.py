"""Throughput of batched Qgemul by instance shape (config-3 types): which kernel the library picks and what it delivers."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qublas_b200 import ops
from qublas_b200.qtypes import QType


def main():
    a = QType(7, 8); mul = QType(14, 8, True, 4, 0); c = QType(17, 4, True, 0, 0)
    shapes = [(16, 16, 16), (32, 32, 32), (32, 32, 64), (64, 64, 64), (64, 64, 128), (128, 128, 64), (128, 64, 32), (64, 64, 48), (48, 48, 48), (8, 8, 64), (64, 1, 64), (20, 12, 64)]
    for (M, N, K) in shapes:
        lv = [QType(15 + l, max(8 - l, 2), True, 0, 0) for l in range(max(1, (K - 1).bit_length()))]
        plan = ops.make_gemul_plan(a, a, c, M, N, K, mul=mul, add=lv)
        words = M * K + K * N + M * N
        batch = max(64, min(1 << 20, (1 << 28) // words))
        A = torch.empty(batch * M * K, dtype=torch.int32, device="cuda"); B = torch.empty(batch * K * N, dtype=torch.int32, device="cuda")
        ops.synth_fill(A, a.raw_min, a.raw_max, 1, 0); ops.synth_fill(B, a.raw_min, a.raw_max, 1, 1)
        Cd = torch.empty(batch * M * N, dtype=torch.int32, device="cuda")
        for _ in range(2):
            ops.Qgemul(Cd, A, B, plan, batch=batch)
        info = ops.plan_info_gemul(plan)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            ops.Qgemul(Cd, A, B, plan, batch=batch)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(json.dumps({"shape": [M, N, K], "batch": batch, "kernel": info.kernel.decode(), "gmacs": round(batch * M * N * K / ms / 1e6, 1),
                          "gbs": round(words * 4 * batch / ms / 1e6, 1)}))
        del A, B, Cd


if __name__ == "__main__":
    main()
