"""Host<->device link ceiling of this box with N GPUs copying AT THE SAME TIME (python tools/pcie_probe.py --gpus N):
per direction alone, both directions at once, with the chunk sizes and 2:1 / 1:1 byte ratios of the library's host
pipeline.  One process per GPU (like bench.py under torchrun), pinned buffers, a barrier before every timed phase; the
result is what bounds the `e2e` figure of bench.py and its 1 -> 8 GPU curve."""
import argparse
import json
import multiprocessing as mp
import os
import time


def worker(rank, world, args, q, barrier):
    import torch
    torch.cuda.set_device(rank)
    if args.numa:
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(rank))
        except Exception:
            pass
    dev = torch.device("cuda", rank)
    n = args.mib << 18                                   # int32 words
    h_in = torch.empty(n, dtype=torch.int32).pin_memory(); h_in.fill_(rank + 1)
    h_out = torch.empty(n, dtype=torch.int32).pin_memory(); h_out.fill_(0)
    d_in = torch.empty(n, dtype=torch.int32, device=dev)
    d_out = torch.zeros(n, dtype=torch.int32, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    chunk = args.chunk_mib << 18

    def h2d(frac=1.0):
        with torch.cuda.stream(s1):
            for o in range(0, int(n * frac), chunk):
                d_in[o:o + chunk].copy_(h_in[o:o + chunk], non_blocking=True)

    def d2h(frac=1.0):
        with torch.cuda.stream(s2):
            for o in range(0, int(n * frac), chunk):
                h_out[o:o + chunk].copy_(d_out[o:o + chunk], non_blocking=True)

    def timed(fn, reps=4):
        fn(); torch.cuda.synchronize()
        barrier.wait()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        barrier.wait()
        return dt

    gb = n * 4 / 1e9
    res = {"rank": rank,
           "h2d": gb / timed(h2d), "d2h": gb / timed(d2h),
           "duplex_1to1_each": gb / timed(lambda: (h2d(), d2h())),
           "duplex_2to1_h2d": gb / timed(lambda: (h2d(), d2h(0.5)))}
    q.put(res)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--mib", type=int, default=1024, help="MiB per direction per GPU and repetition")
    ap.add_argument("--chunk-mib", type=int, default=32, help="copy granularity (the library pipelines 12..96 MiB chunks)")
    ap.add_argument("--numa", action="store_true", help="bind every process to the CPUs next to its GPU before allocating")
    args = ap.parse_args()
    ctx = mp.get_context("spawn")
    q = ctx.Queue(); barrier = ctx.Barrier(args.gpus)
    ps = [ctx.Process(target=worker, args=(r, args.gpus, args, q, barrier)) for r in range(args.gpus)]
    for p in ps:
        p.start()
    rows = [q.get() for _ in ps]
    for p in ps:
        p.join()
    out = {"gpus": args.gpus, "numa_bind": args.numa, "chunk_mib": args.chunk_mib, "host_cpus": os.cpu_count()}
    for k in ("h2d", "d2h", "duplex_1to1_each", "duplex_2to1_h2d"):
        vals = [r[k] for r in rows]
        out[k + "_gbs_per_gpu_min"] = min(vals); out[k + "_gbs_sum"] = sum(vals)
    out["duplex_1to1_total_gbs_sum"] = 2 * out["duplex_1to1_each_gbs_sum"]
    out["duplex_2to1_total_gbs_sum"] = 1.5 * out["duplex_2to1_h2d_gbs_sum"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
