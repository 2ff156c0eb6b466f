"""Host-buffer (end-to-end) rate of the C3 Qgemul as a function of call size; run once per QB200_PIPE_MB value."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qublas_b200 import ops
from qublas_b200.qtypes import QType

def main():
    a = QType(7, 8); mul = QType(14, 8, True, 4, 0); lv = [QType(15 + l, 8 - l, True, 0, 0) for l in range(6)]; c = QType(17, 4, True, 0, 0)
    plan = ops.make_gemul_plan(a, a, c, 64, 64, 64, mul=mul, add=lv)
    out = {}
    for chunk in (8192, 32768):
        A = torch.empty(chunk * 4096, dtype=torch.int32).pin_memory(); B = torch.empty_like(A).pin_memory(); C = torch.empty_like(A).pin_memory()
        dA = torch.empty(chunk * 4096, dtype=torch.int32, device="cuda")
        ops.synth_fill(dA, a.raw_min, a.raw_max, 1, 0); A.copy_(dA); ops.synth_fill(dA, a.raw_min, a.raw_max, 1, 1); B.copy_(dA)
        An, Bn, Cn = A.numpy(), B.numpy(), C.numpy()
        ops.Qgemul(Cn, An, Bn, plan, batch=chunk)
        reps = max(2, 131072 // chunk)
        t0 = time.perf_counter()
        for _ in range(reps):
            ops.Qgemul(Cn, An, Bn, plan, batch=chunk)
        dt = (time.perf_counter() - t0) / reps
        out[chunk] = {"gmacs": chunk * 262144 / dt / 1e9, "h2d_gbs": chunk * 32768 / dt / 1e9}
        del A, B, C
    print(json.dumps({"pipe_mb": os.environ.get("QB200_PIPE_MB", "48"), "by_call_size": out}))

if __name__ == "__main__":
    main()
