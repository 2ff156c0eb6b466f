import json, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qublas_b200 import ops
from qublas_b200.qtypes import QType
def timed(fn, reps=5):
    fn(); fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
n = 1 << 26
t = QType(7, 8, True, 4, 0)
x = torch.randn(n, dtype=torch.float64, device="cuda") * 50
ms = timed(lambda: ops.quantize(x, t)); print(json.dumps({"op": "quantize_f64", "gbs": round(n * 12 / ms / 1e6, 1)}))
w = ops.quantize(x, t)
ms = timed(lambda: ops.dequantize(w, t)); print(json.dumps({"op": "dequantize_f64", "gbs": round(n * 12 / ms / 1e6, 1)}))
ms = timed(lambda: ops.error_stats(w, x, t)); print(json.dumps({"op": "error_stats", "gbs": round(n * 12 / ms / 1e6, 1)}))
