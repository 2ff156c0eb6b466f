"""Host<->device link ceiling on this box: pinned H2D alone, D2H alone, and both directions at once.
The e2e numbers of bench.py (host buffers through the C ABI) are bounded by these."""
import json
import time

import torch


def main():
    dev = torch.device("cuda:0")
    n = 1 << 28                                      # 1 GiB of int32
    h_in = torch.empty(n, dtype=torch.int32).pin_memory()
    h_out = torch.empty(n, dtype=torch.int32).pin_memory()
    d_in = torch.empty(n, dtype=torch.int32, device=dev)
    d_out = torch.zeros(n, dtype=torch.int32, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn, reps=5):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    def h2d():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    def both():
        h2d(); d2h()

    gb = n * 4 / 1e9
    t_h2d, t_d2h, t_both = timed(h2d), timed(d2h), timed(both)
    print(json.dumps({"h2d_gbs": gb / t_h2d, "d2h_gbs": gb / t_d2h, "duplex_each_gbs": gb / t_both, "duplex_total_gbs": 2 * gb / t_both}))


if __name__ == "__main__":
    main()
