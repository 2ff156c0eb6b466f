"""Element-wise kernels (Qmul/Qadd on tensors, TF / Basic complex products, convert, quantize): GB/s of algorithmic traffic."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qublas_b200 import ops
from qublas_b200.qtypes import QType


def timed(fn, reps=5):
    fn(); fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n = 1 << 26
    t = QType(7, 8); to = QType(10, 6, True, 4, 0)
    a = torch.empty(n, dtype=torch.int32, device="cuda"); b = torch.empty_like(a)
    ops.synth_fill(a, t.raw_min, t.raw_max, 1, 0); ops.synth_fill(b, t.raw_min, t.raw_max, 1, 1)
    for op in ("mul", "add"):
        ms = timed(lambda: ops.elementwise(op, a, b, t, t, to=to))
        print(json.dumps({"op": "elementwise " + op, "n": n, "gelems": round(n / ms / 1e6, 1), "gbs": round(n * 12 / ms / 1e6, 1)}))
    ms = timed(lambda: ops.convert(a, t, to))
    print(json.dumps({"op": "convert", "n": n, "gelems": round(n / ms / 1e6, 1), "gbs": round(n * 8 / ms / 1e6, 1)}))
    c = QType(3, 12)
    P = QType(6, 12); S = QType(7, 12)
    tags = {"abT": QType(4, 12), "cdT": QType(4, 11), "abcT": P, "cdbT": P, "badT": P, "ABT": S, "BCT": S}
    for product, tg in (("tf", tags), ("basic", None)):
        ms = timed(lambda: ops.tfcmul(a, b, c, c, c, c, tg, product=product))
        print(json.dumps({"op": "complex product " + product, "n": n // 2, "gelems": round(n / 2 / ms / 1e6, 1), "gbs": round(n // 2 * 24 / ms / 1e6, 1)}))


if __name__ == "__main__":
    main()
