"""Qreduce throughput by vector length (config-2 types): kernel chosen and fraction of HBM bandwidth."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qublas_b200 import ops
from qublas_b200.qtypes import QType, TREE_VECTOR, TREE_VARIADIC


def main():
    elem = QType(8, 8)
    for form in (TREE_VECTOR, TREE_VARIADIC):
        for n in (64, 1000, 4096, 65536, 100000, 1 << 20, 3 << 20, (1 << 24) - 1, 1 << 24):
            nlev = max(1, (n - 1).bit_length())
            lv = [QType(9 + l, 8 - l, True, 0, 0) for l in range(8)] + [QType(17 + (l - 6) // 2, 0, True, 0, 0) for l in range(8, nlev)]
            lv = lv[:nlev]
            batch = max(1, (1 << 26) // n)
            x = torch.empty(batch * n, dtype=torch.int32, device="cuda")
            ops.synth_fill(x, elem.raw_min, elem.raw_max, 1, 0)
            plan = ops.make_reduce_plan(elem, lv, n, form=form)
            out = torch.empty(batch, dtype=torch.int32, device="cuda")
            try:
                for _ in range(2):
                    ops.Qreduce(x, elem, lv, length=n, batch=batch, plan=plan, out=out)
            except Exception as e:
                print(json.dumps({"len": n, "form": form, "error": str(e)[:100]})); continue
            info = ops.plan_info_reduce(plan)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                ops.Qreduce(x, elem, lv, length=n, batch=batch, plan=plan, out=out)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            print(json.dumps({"len": n, "form": "vector" if form == TREE_VECTOR else "variadic", "batch": batch, "kernel": info.kernel.decode(),
                              "gelems": round(batch * n / ms / 1e6, 1), "gbs": round(batch * n * 4 / ms / 1e6, 1)}))
            del x


if __name__ == "__main__":
    main()
