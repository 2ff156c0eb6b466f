"""Kernel experiments on BASELINE config 3 (device-resident): GMAC/s of the kernel the library picks under the current
environment (QB200_NO_AOT=1 QB200_JIT_DEFINES="-D..." select NVRTC builds of the same source with other knobs), checked
against the oracle on a few instances.  python tools/c3_variants.py [batch]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle  # noqa: E402
from qublas_b200 import ops  # noqa: E402
from qublas_b200.qtypes import QType, RND_CONV, RND_POS_INF, SAT_TCPL  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
    adversarial = len(sys.argv) > 2 and sys.argv[2] == "adv"
    a = QType(7, 8); mul = QType(14, 8, True, RND_CONV, SAT_TCPL)
    lv = [QType(15 + l, 8 - l, True, RND_POS_INF, SAT_TCPL) for l in range(6)]
    c = QType(17, 4, True, RND_POS_INF, SAT_TCPL)
    plan = ops.make_gemul_plan(a, a, c, 64, 64, 64, mul=mul, add=lv)
    A = torch.empty(batch * 4096, dtype=torch.int32, device="cuda"); B = torch.empty_like(A); Cd = torch.empty_like(A)
    ops.synth_fill(A, a.raw_min, a.raw_max, 1, 0); ops.synth_fill(B, a.raw_min, a.raw_max, 1, 1)
    if adversarial:                                   # every instance holds the minimum in A and in B: the clamped variant runs
        A.view(batch, 4096)[:, 7] = a.raw_min; B.view(batch, 4096)[:, 9] = a.raw_min
    ops.Qgemul(Cd, A, B, plan, batch=batch)
    torch.cuda.synchronize()
    n = 24
    want = oracle.gemul_batched(plan, A[:n * 4096].cpu().numpy(), B[:n * 4096].cpu().numpy(), n)
    ok = bool(np.array_equal(Cd[:n * 4096].cpu().numpy(), want))
    for _ in range(3):
        ops.Qgemul(Cd, A, B, plan, batch=batch)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        ops.Qgemul(Cd, A, B, plan, batch=batch)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"defines": os.environ.get("QB200_JIT_DEFINES", ""), "aot": not os.environ.get("QB200_NO_AOT"), "adversarial": adversarial,
                      "kernel": ops.plan_info_gemul(plan).kernel.decode(), "bit_exact": ok, "ms": ms, "gmacs": batch * 262144 / ms / 1e6}))


if __name__ == "__main__":
    main()
