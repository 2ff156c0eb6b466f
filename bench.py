#!/usr/bin/env python3
"""bench.py — headline metric of BASELINE.json: bit-exact quantized GMAC/s, batched Qgemul.

Workload (config.workload): BASELINE config 3 — `batch` independent 64x64x64 Qgemul instances per
GPU, 16-bit signed operands Qu<7,8>, QgemulMulArgs Qu<14,8,RND::CONV,SAT::TCPL>, QgemulAddArgs
TypeList<Qu<15+l,8-l,RND::POS_INF,SAT::TCPL>, l=0..5>, C Qu<17,4,RND::POS_INF,SAT::TCPL>
(SURVEY.md 8(d)); raw int32 words in the reference's layout, synthetic counter-RNG inputs.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

N > 1: launched under torchrun, one rank per GPU, instances sharded by batch (no data-path
collective; weak scaling: B instances per GPU).  A step = one pass of the hot path over the
rank's B instances.  `value` = device-timed throughput with inputs resident in HBM; `e2e` = the
same pass through the public API with HOST buffers (pinned), H2D/D2H inside the timed region.
`--impl reference` times the reference's own CPU code (oracle/_ref, compiled from the unmodified
header) on the box's host cores.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M = N = K = 64
MACS_PER_INSTANCE = M * N * K
WORDS = M * K                      # words per matrix (all three are 64x64)
METRIC = "bit-exact quantized GMAC/s, batched Qgemul (64x64x64, 16-bit operands, per-product + per-level quantization)"
METRICS = {"c3": METRIC,
           "c2": "bit-exact Qreduce throughput, 2^24-element Qu<8,8> vectors, 24-level TypeList (RND::POS_INF)",
           "c4": "bit-exact TFComplexMul (3-mult/5-add) complex mat-vec throughput, batched 256x256",
           "c5": "bit-exact Qgemul 8192^3, int8-range operands, lossless accumulators (tcgen05 kind::i8), incl. int32->int8 pack"}


def c3_types():
    from qublas_b200.qtypes import QType, RND_CONV, RND_POS_INF, SAT_TCPL
    a = QType(7, 8)
    mul = QType(14, 8, True, RND_CONV, SAT_TCPL)
    lv = [QType(15 + l, 8 - l, True, RND_POS_INF, SAT_TCPL) for l in range(6)]
    c = QType(17, 4, True, RND_POS_INF, SAT_TCPL)
    return a, mul, lv, c


# ------------------------------------------------------------------------------------------------
# CPU baseline: the reference's own implementation (oracle/_ref) or, if absent, the oracle port
# ------------------------------------------------------------------------------------------------
WORKLOADS = {
    "c3": "BASELINE config 3: batched 64x64x64 Qgemul, 16-bit signed operands Qu<7,8>, MulArgs Qu<14,8,RND::CONV,SAT::TCPL>, "
          "6-level TypeList tree RND::POS_INF, C Qu<17,4>",
    "c2": "BASELINE config 2 (C2b): Qreduce of 2^24 x Qu<8,8>, 24-level TypeList (RND::POS_INF), batch of 16 vectors (1 GiB, > L2)",
    "c4": "BASELINE config 4: batched 256x256 complex mat-vec, TFComplexMul with mixed abT/cdT types, 8-level RND::CONV tree",
    "c5": "BASELINE config 5: 8192^3 Qgemul, A,B Qu<7,0> raws in [-128,127], MulArgs Qu<15,0>, AddArgs Qu<30,0>, C Qu<30,0>",
}


_REF_PICK = None


def ref_lib_path():
    """oracle/_ref = the unmodified reference header compiled in the container that has it (oracle/Makefile).  Two builds
    travel: x86-64-v4 (AVX-512; what -march=native gives on the Xeon hosts) and x86-64-v3.  Both are timed once on a few
    config-3 instances and the faster one that this CPU can run is used (the reference arm gets its best build)."""
    global _REF_PICK
    if _REF_PICK is not None:
        return _REF_PICK
    if os.environ.get("QB_REF_SO"):
        _REF_PICK = (os.environ["QB_REF_SO"], os.environ.get("QB_REF_BUILD", "?"))
        return _REF_PICK
    d = os.path.join(ROOT, "oracle", "_ref")
    try:
        flags = set(open("/proc/cpuinfo").read().split("flags")[1].split("\n")[0].split())
    except Exception:
        flags = set()
    cands = []
    v4 = os.path.join(d, "libqref_bench_v4.so")
    if {"avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"} <= flags and os.path.exists(v4):
        cands.append((v4, "-O3 -march=x86-64-v4"))
    v3 = os.path.join(d, "libqref_bench.so")
    if os.path.exists(v3):
        cands.append((v3, "-O3 -march=x86-64-v3"))
    if not cands:
        _REF_PICK = (None, None)
        return _REF_PICK
    if len(cands) > 1:
        import numpy as np
        from oracle import oracle
        a = c3_types()[0]
        A = oracle.synth_fill(4 * WORDS, 4, a.raw_min, a.raw_max, 1, 0); B = oracle.synth_fill(4 * WORDS, 4, a.raw_min, a.raw_max, 1, 1)
        Cc = np.zeros(4 * WORDS, dtype=np.int32)
        vp = ctypes.c_void_p
        best = None
        for so, build in cands:
            lib = ctypes.CDLL(so)
            lib.qref_c3_gemul.argtypes = [vp, vp, vp, ctypes.c_int]
            dt = 1e9
            for _ in range(3):
                t0 = time.perf_counter(); lib.qref_c3_gemul(A.ctypes.data_as(vp), B.ctypes.data_as(vp), Cc.ctypes.data_as(vp), 4)
                dt = min(dt, time.perf_counter() - t0)
            if best is None or dt < best[0]:
                best = (dt, so, build)
        _REF_PICK = (best[1], best[2])
    else:
        _REF_PICK = cands[0]
    os.environ["QB_REF_SO"], os.environ["QB_REF_BUILD"] = _REF_PICK
    return _REF_PICK


def c4_types():
    from qublas_b200.qtypes import QType, RND_CONV, SAT_TCPL
    t = QType(3, 12); P = QType(6, 12); S = QType(7, 12)
    tags = {"abT": QType(4, 12), "cdT": QType(4, 11), "abcT": P, "cdbT": P, "badT": P, "ABT": S, "BCT": S}
    lv = [QType(8 + l, 12 - l, True, RND_CONV, SAT_TCPL) for l in range(8)]
    return t, tags, lv


def c2_types():
    from qublas_b200.qtypes import QType, RND_POS_INF, SAT_TCPL
    elem = QType(8, 8)
    lv = [QType(9 + l, 8 - l, True, RND_POS_INF, SAT_TCPL) for l in range(8)]
    lv += [QType(17 + (l - 6) // 2, 0, True, RND_POS_INF, SAT_TCPL) for l in range(8, 24)]      # C2b: 24-entry list
    return elem, lv


def _cpu_worker(args):
    """One host core: `count` units of workload `w` starting at unit `first` of the synthetic stream, `reps` times.
    Returns (seconds, checksum).  kind "reference" = oracle/_ref (the reference's own code), "port" = the C oracle."""
    w, kind, first, count, reps = args
    import numpy as np
    from oracle import oracle
    vp = ctypes.c_void_p
    lib = None
    if kind == "reference":
        lib = ctypes.CDLL(ref_lib_path()[0])
    if w == "c3":
        a, mul, lv, c = c3_types()
        A = oracle.synth_fill(count * WORDS, 4, a.raw_min, a.raw_max, 1, 0, first * WORDS)
        B = oracle.synth_fill(count * WORDS, 4, a.raw_min, a.raw_max, 1, 1, first * WORDS)
        Cc = np.zeros(count * WORDS, dtype=np.int32)
        if lib is not None:
            lib.qref_c3_gemul.argtypes = [vp, vp, vp, ctypes.c_int]
            t0 = time.perf_counter()
            for _ in range(reps):
                lib.qref_c3_gemul(A.ctypes.data_as(vp), B.ctypes.data_as(vp), Cc.ctypes.data_as(vp), count)
            dt = time.perf_counter() - t0
        else:
            from qublas_b200 import ops
            plan = ops.make_gemul_plan(a, a, c, M, N, K, mul=mul, add=lv)
            t0 = time.perf_counter()
            for _ in range(reps):
                Cc = oracle.gemul_batched(plan, A, B, count)
            dt = time.perf_counter() - t0
        return dt, int(Cc[:64].astype(np.int64).sum())
    if w == "c4":
        from qublas_b200 import ops
        from qublas_b200.qtypes import TREE_VARIADIC
        t, tags, lv = c4_types()
        H = oracle.synth_fill(count * 256 * 256 * 2, 4, -(1 << 14), (1 << 14) - 1, 1, 0, first * 256 * 256 * 2)
        x = oracle.synth_fill(count * 512, 4, -(1 << 14), (1 << 14) - 1, 1, 1, first * 512)
        y = np.zeros(count * 512, dtype=np.int32)
        if lib is not None:
            lib.qref_c4_matvec.argtypes = [vp, vp, vp, ctypes.c_int]
            t0 = time.perf_counter()
            for _ in range(reps):
                lib.qref_c4_matvec(H.ctypes.data_as(vp), x.ctypes.data_as(vp), y.ctypes.data_as(vp), count)
            dt = time.perf_counter() - t0
        else:
            plan = ops.make_cgemul_plan(t, t, t, t, lv[-1], lv[-1], 256, 1, 256, tf_tags=tags, add_re=lv, form=TREE_VARIADIC)
            t0 = time.perf_counter()
            for _ in range(reps):
                y = oracle.cgemul_batched(plan, H, x, count)
            dt = time.perf_counter() - t0
        return dt, int(y[:64].astype(np.int64).sum())
    if w == "c2":                      # a unit = one 2^24-element vector
        from qublas_b200 import ops
        elem, lv = c2_types()
        n = 1 << 24
        out = np.zeros(1, dtype=np.int32)
        dt = 0.0
        for v in range(count):
            xv = oracle.synth_fill(n, 4, elem.raw_min, elem.raw_max, 1, 0, (first + v) * n)
            t0 = time.perf_counter()
            for _ in range(reps):
                if lib is not None:
                    lib.qref_c2_chain.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp]
                    assert lib.qref_c2_chain(xv.ctypes.data_as(vp), 1, 0, out.ctypes.data_as(vp), None) == 0
                else:
                    out, _ = oracle.reduce_batched(ops.make_reduce_plan(elem, lv, n), xv, 1)
            dt += time.perf_counter() - t0
        return dt, int(out[0])
    raise ValueError(w)


# units and work per unit of each workload (metric numerator)
UNIT_WORK = {"c3": MACS_PER_INSTANCE, "c4": 256 * 256, "c2": 1 << 24}
UNIT_NAME = {"c3": "instances", "c4": "instances", "c2": "vectors"}
UNIT_OUT = {"c3": "GMAC/s", "c4": "G complex-MAC/s", "c2": "Gelem/s"}


def cpu_kind():
    return "reference" if ref_lib_path()[0] else "port"


def cpu_rate(w, per_core_units, cores, kind):
    """All-core throughput (G work-items/s) on a bounded sample: every core runs `per_core_units` distinct units of the
    workload (fork, one process per core -- the reference is not thread-safe)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    jobs = [(w, kind, r * per_core_units, per_core_units, 1) for r in range(cores)]
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    busy = max(r[0] for r in res)
    total = per_core_units * cores * UNIT_WORK[w]
    return total / busy / 1e9, wall


def cpu_baseline(w, seconds):
    """cpu_baseline object: the reference's own CPU path on all host cores, ~`seconds` of work per core."""
    from oracle import oracle
    oracle.build()
    kind = cpu_kind()
    cores = os.cpu_count() or 1
    probe = {"c3": 4 if kind == "reference" else 1, "c4": 1, "c2": 1}[w]
    t1, _ = _cpu_worker((w, kind, 0, probe, 1))
    per_core = max(1, int(seconds / (t1 / probe)))
    g, wall = cpu_rate(w, per_core, cores, kind)
    return {"value": g, "unit": UNIT_OUT[w], "cores": cores, "kind": kind,
            "build": ref_lib_path()[1] if kind == "reference" else "gcc -O2 (C oracle)",
            "sample": "%d %s/core x %d cores (%.1f s wall) of the same synthetic stream" % (per_core, UNIT_NAME[w], cores, wall)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    oracle.build()
    w = args.workload
    if w not in UNIT_WORK:
        print(json.dumps({"impl": "reference", "unavailable": "no CPU reference arm for workload %s (the header cannot instantiate it)" % w}), flush=True)
        return
    kind = cpu_kind()
    cores = os.cpu_count() or 1
    # calibrate on one core, then size a step for ~2.5 s of all-core work
    probe = {"c3": 4 if kind == "reference" else 1, "c4": 1, "c2": 1}[w]
    t, _ = _cpu_worker((w, kind, 0, probe, 1))
    per_core = max(1, int(2.5 / (t / probe)))
    vals = []
    for step in range(args.warmup + args.steps):
        g, _ = cpu_rate(w, per_core, cores, kind)
        if step >= args.warmup:
            vals.append(g)
    value = statistics.mean(vals)
    sample = "%d %s/core x %d cores per step (bounded sample of the workload)" % (per_core, UNIT_NAME[w], cores)
    line = {
        "impl": "reference", "metric": METRICS[w], "value": value, "unit": UNIT_OUT[w], "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * per_core * cores * UNIT_WORK[w] / (value * 1e9),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config(w, args),
        "cpu_baseline": {"value": value, "unit": UNIT_OUT[w], "cores": cores, "kind": kind,
                         "build": ref_lib_path()[1] if kind == "reference" else "gcc -O2 (C oracle)", "sample": sample},
        "e2e": {"value": value, "unit": UNIT_OUT[w], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(w, args):
    """`config`: printed identically by both arms (the driver compares it); what is specific to one arm goes to `details`."""
    cfg = {"workload": WORKLOADS[w]}
    if w == "c3":
        cfg["batch_per_gpu"] = args.batch
        cfg["l2"] = "no flush needed: the inputs of a step (%.1f GB per GPU) are far larger than the 126 MB L2" % (2 * args.batch * WORDS * 4 / 1e9)
    return cfg


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.nvml, self.stop_flag = None, False

    def _nvml_loop(self):
        n, h = self.nvml
        names = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM); mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
                pw = n.nvmlDeviceGetPowerUsage(h) / 1000.0; mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((float(sm), float(mx), pw, [v for k, v in names.items() if mask & k]))
            except Exception:
                break
            time.sleep(0.005)

    def sample_now(self):
        """One sample taken by the calling thread (after the timed steps have been enqueued and while the GPU still works on
        them): the region is sampled even on a box where the sampling thread's first NVML calls are slow."""
        if not self.nvml:
            return
        n, h = self.nvml
        names = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}
        try:
            sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM); mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
            pw = n.nvmlDeviceGetPowerUsage(h) / 1000.0; mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            self.rows.append((float(sm), float(mx), pw, [v for k, v in names.items() if mask & k]))
        except Exception:
            pass

    def start(self):
        # NVML in-process (a sample every ~5 ms, so even a 100 ms timed region gets a median); nvidia-smi -lms as the fallback
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = (pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index))
            self.t = threading.Thread(target=self._nvml_loop, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.nvml:
            self.stop_flag = True
            self.t.join(timeout=2)
            sm = [r[0] for r in self.rows]; mx = [r[1] for r in self.rows]; pw = [r[2] for r in self.rows]
            reasons = sorted({x for r in self.rows for x in r[3]})
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                    "power_w_max": max(pw) if pw else None, "samples": len(sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower().startswith("active")})
        pw = [float(r[3]) for r in self.rows if len(r) >= 8 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(sm)}


# algorithmic bytes of the launch that each committed `ncu --set full` capture profiled (tools/prof.py)
NCU_CAPTURE_ALG_BYTES = {"c3": 32768 * 3 * 4096 * 4, "c2": 16 * (1 << 24) * 4 + 16 * 4, "c4": 4096 * (256 * 256 + 256 + 256) * 8,
                         "c5": 2 * 8192 * 8192 + 8192 * 8192 * 4}


def ncu_traffic(tag, alg_bytes_per_launch):
    """DRAM bytes per launch of the dominant kernel: dram__bytes_read.sum + dram__bytes_write.sum of the committed
    `ncu --set full` capture (profiles/r1_<tag>_ncu_summary.json), scaled from the capture's launch size to this
    run's by the ratio of algorithmic bytes (the kernels stream every byte once, so traffic is linear in the batch)."""
    try:
        path = os.path.join(ROOT, "profiles", "r2_%s_ncu_summary.json" % tag)
        if not os.path.exists(path):
            path = os.path.join(ROOT, "profiles", "r1_%s_ncu_summary.json" % tag)
        prof = json.load(open(path))[0]
        return float(prof["traffic_bytes"]) * alg_bytes_per_launch / NCU_CAPTURE_ALG_BYTES[tag]
    except Exception:
        return None


def roofline_alu(lib, gmacs, device, sm_mhz=None):
    """The headline kernel is bound by instruction issue on the INT32/FP32 pipes, not by HBM.  Three yardsticks:
      frac      achieved thread-instruction rate (instructions per MAC of the committed ncu capture of THIS round's kernel)
                over the issue peak measured live by qb200_measure_int_peak (IMAD + IADD3/LOP3 mix on all SMs);
      frac_hw   the same over the hardware's issue rate, SMs x 4 schedulers x 32 lanes x the SM clock sampled in this run;
      frac_mix  GMAC/s over the rate the SAME instruction mix reaches from registers alone (no shared memory, no barriers:
                tools/mix_probe.cu, profiles/r2_mix_probe.jsonl) -- the ceiling of this formulation of the arithmetic."""
    from qublas_b200 import abi
    out = {"bound": "issue (INT32 ALU + FP32 FMA pipes)", "unit": "T thread-instr/s", "scope": "per GPU"}
    try:
        name = "r2_c3_ncu_summary.json" if os.path.exists(os.path.join(ROOT, "profiles", "r2_c3_ncu_summary.json")) else "r1_c3_ncu_summary.json"
        prof = json.load(open(os.path.join(ROOT, "profiles", name)))[0]
        ipm = float(prof["thread_instructions_per_unit"])
        out.update({"instr_per_mac": ipm, "ncu_pipe_alu_pct": prof.get("pipe_alu_pct"), "ncu_pipe_fma_pct": prof.get("pipe_fma_pct"),
                    "ncu_issue_active_pct": prof.get("issue_active_pct"), "source": "profiles/" + name})
    except Exception:
        return out
    out["achieved"] = gmacs * 1e9 * ipm / 1e12
    t = ctypes.c_double()
    ex = abi.qb_exec(abi.QB_MEM_DEVICE, device, None, 0, 0)
    if lib.qb200_measure_int_peak(0.25, ctypes.byref(t), ctypes.byref(ex)) == 0 and t.value > 0:
        out["peak"] = t.value
        out["frac"] = out["achieved"] / out["peak"]
        out["peak_source"] = "qb200_measure_int_peak (IMAD + IADD3/LOP3 mix, all SMs), this run"
    try:
        import torch
        sms = torch.cuda.get_device_properties(device).multi_processor_count
        if sm_mhz:
            out["peak_hw"] = sms * 4 * 32 * sm_mhz * 1e6 / 1e12
            out["frac_hw"] = out["achieved"] / out["peak_hw"]
            out["peak_hw_source"] = "%d SMs x 4 schedulers x 32 lanes x %.0f MHz (median SM clock of the timed region)" % (sms, sm_mhz)
    except Exception:
        pass
    try:
        best = 0.0
        for ln in open(os.path.join(ROOT, "profiles", "r2_mix_probe.jsonl")):
            if ln.startswith("{") and "tmacs_equiv" in ln:
                best = max(best, float(json.loads(ln)["tmacs_equiv"]))
        if best > 0:
            out["mix_ceiling_gmacs"] = best * 1e3
            out["frac_mix"] = gmacs / (best * 1e3)
            out["mix_source"] = "profiles/r2_mix_probe.jsonl (tools/mix_probe.cu: the kernel's FFMA/IADD3/SHF/LEA mix from registers only, best variant)"
    except Exception:
        pass
    return out


# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local):
    """Multi-GPU end-to-end runs move ~75 GB/s per GPU through host memory: keep each rank's threads (and so the
    first-touch placement of its pinned buffers) on the CPUs next to its GPU.  Best effort; returns what was done."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return "cpus %d" % len(os.sched_getaffinity(0))
    except Exception as e:            # no NVML, a single NUMA node, a restricted container: nothing to do
        return "unchanged (%s)" % type(e).__name__


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from qublas_b200 import abi, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    lib = abi.load()
    if lib.qb200_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = bind_to_gpu_numa_node(local) if world > 1 else "unchanged (single GPU)"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    a, mul, lv, c = c3_types()
    plan = ops.make_gemul_plan(a, a, c, M, N, K, mul=mul, add=lv)
    info = ops.plan_info_gemul(plan)
    Bn = args.batch
    first, _ = rank * Bn, Bn                         # weak scaling: every rank owns B instances of one global stream
    A = torch.empty(Bn * WORDS, dtype=torch.int32, device=dev)
    B = torch.empty(Bn * WORDS, dtype=torch.int32, device=dev)
    Cd = torch.empty(Bn * WORDS, dtype=torch.int32, device=dev)
    ops.synth_fill(A, a.raw_min, a.raw_max, 1, 0, first * WORDS)
    ops.synth_fill(B, a.raw_min, a.raw_max, 1, 1, first * WORDS)

    # ---- parity check (not timed): >= 512 full instances of this rank -- the first, the last, and evenly strided ones in
    # between -- bit for bit against the oracle ----
    from oracle import oracle
    ops.Qgemul(Cd, A, B, plan, batch=Bn)
    torch.cuda.synchronize()
    n_chk = min(Bn, args.check_instances)
    picks = sorted({0, Bn - 1} | {int(i * (Bn - 1) / max(n_chk - 1, 1)) for i in range(n_chk)})
    idx = torch.tensor(picks, dtype=torch.int64, device=dev)
    gA = A.view(Bn, WORDS).index_select(0, idx).cpu().numpy(); gB = B.view(Bn, WORDS).index_select(0, idx).cpu().numpy()
    gC = Cd.view(Bn, WORDS).index_select(0, idx).cpu().numpy()

    def _check(lo_hi):
        lo, hi = lo_hi
        want = oracle.gemul_batched(plan, gA[lo:hi].reshape(-1), gB[lo:hi].reshape(-1), hi - lo)      # ctypes releases the GIL
        return np.array_equal(gC[lo:hi].reshape(-1), want)

    from concurrent.futures import ThreadPoolExecutor
    nthr = max(1, min(16, (os.cpu_count() or 1) // world))
    cuts = [(i * len(picks) // nthr, (i + 1) * len(picks) // nthr) for i in range(nthr)]
    with ThreadPoolExecutor(nthr) as pool:
        oks = list(pool.map(_check, [c for c in cuts if c[1] > c[0]]))
    if not all(oks):
        raise SystemExit("bench.py: GPU result differs from the oracle on rank %d" % rank)
    checked = "%d full instances per rank (first, last, evenly strided) vs the oracle, all ranks" % len(picks)

    # ---- device-resident timing: K steps, one C-ABI call per step ----
    for _ in range(args.warmup):
        ops.Qgemul(Cd, A, B, plan, batch=Bn)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.qb200_launch_count_reset()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        ops.Qgemul(Cd, A, B, plan, batch=Bn)
    ev1.record()
    if rank == 0:
        sampler.sample_now(); sampler.sample_now()          # the steps are enqueued, the GPU is still inside the timed region
    barrier()
    launches = int(lib.qb200_launch_count())
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * Bn * MACS_PER_INSTANCE / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the C++ drop-in API (include/QuBLAS.h): qublas_b200/libqb200_e2e.so is user code written
    # against the header -- Qbatch containers in pinned host memory, QgemulBatched<QgemulAddArgs<..>, QgemulMulArgs<..>>(C, A, B)
    # per chunk; the H2D copy of every chunk's operands and the D2H copy of its results happen inside the call ----
    e2e_lib = ctypes.CDLL(os.path.join(ROOT, "qublas_b200", "libqb200_e2e.so"))
    e2e_lib.qbx_last_error.restype = ctypes.c_char_p
    u64 = ctypes.c_uint64
    e2e_lib.qbx_c3_init.argtypes = [ctypes.c_int, u64, ctypes.c_int, u64]
    e2e_lib.qbx_c3_step.argtypes = [ctypes.c_int, u64, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(u64), ctypes.POINTER(u64)]
    e2e_lib.qbx_c3_result.argtypes = [ctypes.c_int, ctypes.c_void_p, u64]
    e2e_lib.qbx_c3_sweep.argtypes = [u64, u64, ctypes.c_int, u64] + [ctypes.POINTER(ctypes.c_double)] * 3
    chunk = min(args.e2e_chunk, Bn)
    e2e_instances = max(chunk, min(Bn, args.e2e_instances) // chunk * chunk)        # whole chunks
    e2e_steps = max(1, min(args.steps, 3))
    want0 = Cd[:WORDS].cpu().numpy()

    def e2e_run(words):
        """words 0: the header's native narrow words (int16 operands, int32 results); 1: the reference's int32 words."""
        if e2e_lib.qbx_c3_init(words, chunk, local, first) != 0:
            raise SystemExit("bench.py: " + e2e_lib.qbx_last_error().decode())
        sec, h2d, d2h = ctypes.c_double(), u64(), u64()

        def step():
            if e2e_lib.qbx_c3_step(words, e2e_instances, ctypes.byref(sec), ctypes.byref(h2d), ctypes.byref(d2h)) != 0:
                raise SystemExit("bench.py: " + e2e_lib.qbx_last_error().decode())
        step()                                                   # warm-up (staging buffers, pipeline streams)
        got = np.zeros(WORDS, dtype=np.int32)
        e2e_lib.qbx_c3_result(words, got.ctypes.data_as(ctypes.c_void_p), WORDS)
        if not np.array_equal(got, want0):
            raise SystemExit("bench.py: C++ host-path result differs from the device-path result")
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step()
        barrier()
        dt = (time.perf_counter() - t0) / e2e_steps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_lib.qbx_c3_free()
        return float(tt.item()), int(h2d.value), int(d2h.value)

    e2e_s, h2d_b, d2h_b = e2e_run(0)
    e2e_value = world * e2e_instances * MACS_PER_INSTANCE / e2e_s / 1e9
    e2e_ref_words = None
    if not args.no_ref_words:
        s32, h32, d32 = e2e_run(1)
        e2e_ref_words = {"value": world * e2e_instances * MACS_PER_INSTANCE / s32 / 1e9, "unit": "GMAC/s", "h2d_bytes_per_step": world * h32,
                         "d2h_bytes_per_step": world * d32, "note": "the same C++ call with Qwords::reference containers (the reference's int32 words): 48 KiB instead of 32 KiB per instance over PCIe"}

    # supplementary: the Monte-Carlo sweep north_star names, device-resident through the same C++ API: inputs generated in
    # HBM, Qgemul, float64 yardstick GEMM and SQNR / max-error statistics on the GPU; 32 bytes per chunk cross PCIe
    sweep = None
    if not args.no_sweep:
        sec, sq, mx = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        n_sw = min(Bn, 4 * chunk)
        for _ in range(2):
            if e2e_lib.qbx_c3_sweep(n_sw, chunk, local, first, ctypes.byref(sec), ctypes.byref(sq), ctypes.byref(mx)) != 0:
                raise SystemExit("bench.py: " + e2e_lib.qbx_last_error().decode())
        sweep = {"value": n_sw * MACS_PER_INSTANCE / sec.value / 1e9, "unit": "GMAC/s", "scope": "rank 0's GPU", "instances": n_sw, "sqnr_db": sq.value, "max_abs_err": mx.value,
                 "note": "synth + Qgemul + float64 reference GEMM + error statistics, all on the device (Qbatch(Qmem::device), QgemulError)"}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        kernel_ms = ms_per_step                       # one kernel launch per step (launches / steps == 1)
        bytes_per_launch = Bn * 3 * WORDS * 4         # A + B read, C written: 48 KiB per instance
        achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "GMAC/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": workload_config("c3", args),
            "details": {"sharding": "batch, no data-path collective", "kernel": info.kernel.decode(),
                        "e2e_host_ring": "2 pinned slots x %d instances, reused cyclically; %d instances per e2e step" % (chunk, e2e_instances),
                        "cpu_affinity": affinity},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": ncu_traffic("c3", bytes_per_launch), "algorithmic_bytes": bytes_per_launch,
                         "traffic_source": "profiles/%s_c3_ncu_summary.json (32768-instance launch), scaled to this launch" % ("r2" if os.path.exists(os.path.join(ROOT, "profiles", "r2_c3_ncu_summary.json")) else "r1"),
                         "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                         "note": "this kernel is bound by the INT32 issue rate, not HBM: see roofline_alu"},
            "roofline_alu": roofline_alu(lib, value / world, local, clocks.get("sm_mhz") if isinstance(clocks, dict) else None),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "GMAC/s", "h2d_bytes_per_step": world * h2d_b, "d2h_bytes_per_step": world * d2h_b,
                    "scope": "whole job (all ranks)",
                    "api": "C++ drop-in header: Qbatch (pinned, native narrow words: int16 operands / int32 results) + QgemulBatched, qublas_b200/cxx/e2e_c3.cpp"},
            "gpu_launches": launches,
            "e2e_reference_words": e2e_ref_words,
            "device_resident_sweep": sweep,
            "bit_exact_checked": checked,
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline("c3", 12.0)
        if world == 1 and not args.no_secondary:
            # the other BASELINE configurations, measured in the same run with the same timing discipline (a few seconds each;
            # `python bench.py --workload c2|c4|c5` prints their full lines, with their own CPU baselines)
            del A, B, Cd
            torch.cuda.empty_cache()
            sec = {}
            for w in ("c2", "c4", "c5"):
                try:
                    sub = argparse.Namespace(**{**vars(args), "workload": w, "no_cpu": True})
                    r = run_secondary(sub, embed=True)
                    sec[w] = {k: r[k] for k in ("metric", "value", "unit", "ms_per_step", "roofline", "clocks", "e2e", "gpu_launches", "int8_operands") if k in r}
                    sec[w]["kernel"] = r["config"].get("kernel")
                except Exception as ex:          # a secondary leg must never cost the headline line
                    sec[w] = {"error": "%s: %s" % (type(ex).__name__, ex)}
                torch.cuda.empty_cache()
            line["secondary"] = sec
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()



# ------------------------------------------------------------------------------------------------
# secondary workloads (BASELINE configs 2, 4, 5): same timing discipline, one JSON line each
# ------------------------------------------------------------------------------------------------
def run_secondary(args, embed=False):
    """One JSON line for BASELINE config 2, 4 or 5; embed=True (single GPU): return the line instead of printing it."""
    import numpy as np
    import torch
    from oracle import oracle
    from qublas_b200 import abi, ops
    from qublas_b200.qtypes import QType, RND_CONV, RND_POS_INF, SAT_TCPL, TREE_VARIADIC, tree_out_type, resolve_tf

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    lib = abi.load()
    if lib.qb200_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not embed:
        dist.init_process_group("nccl", device_id=dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    w = args.workload

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()
        lib.qb200_launch_count_reset()
        sampler = ClockSampler(local); sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # a step of these kernels lasts ~1 ms: repeat the K-step block so that the clock sampler sees the load
        t_probe0 = time.perf_counter(); fn(); torch.cuda.synchronize(); probe = max(time.perf_counter() - t_probe0, 1e-5)
        reps = max(1, int(1.2 / (probe * args.steps)))
        if world > 1:
            rt = torch.tensor([reps], dtype=torch.int64, device=dev); dist.broadcast(rt, 0); reps = int(rt.item())
        lib.qb200_launch_count_reset()
        e0.record()
        for _ in range(args.steps * reps):
            fn()
        e1.record()
        sampler.sample_now()
        torch.cuda.synchronize()
        ms_local = e0.elapsed_time(e1) / (args.steps * reps)
        if world > 1:                                   # device time, max over ranks
            tt = torch.tensor([ms_local], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms_local = float(tt.item())
        return ms_local, int(lib.qb200_launch_count()) // reps, sampler.stop()

    def wall(fn, n=2):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n

    if w == "c2":
        n, batch = 1 << 24, 16
        elem, lv = c2_types()
        x = torch.empty(batch * n, dtype=torch.int32, device=dev)
        ops.synth_fill(x, elem.raw_min, elem.raw_max, 1, 0)
        out, _ = ops.Qreduce(x, elem, lv, length=n, batch=batch)
        want, _ = oracle.reduce_batched(ops.make_reduce_plan(elem, lv, n), x[:n].cpu().numpy(), 1)
        assert int(out.cpu()[0]) == int(want[0]), "Qreduce differs from the oracle"
        rplan = ops.make_reduce_plan(elem, lv, n)
        rout = torch.empty(batch, dtype=torch.int32, device=dev)
        ms, launches, clocks = timed(lambda: ops.Qreduce(x, elem, lv, length=n, batch=batch, plan=rplan, out=rout))
        units = batch * n
        value, unit, metric = world * units / (ms * 1e-3) / 1e9, "Gelem/s", "bit-exact Qreduce throughput, 2^24-element Qu<8,8> vectors, 24-level TypeList (RND::POS_INF)"
        bytes_alg = units * 4 + batch * 4
        roof = {"bound": "hbm", "achieved": bytes_alg / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s"}
        hx = torch.empty(n, dtype=torch.int32).pin_memory(); hx.copy_(x[:n]); hxn = hx.numpy()
        e2e_s = wall(lambda: ops.Qreduce(hxn, elem, lv))
        e2e = {"value": n / e2e_s / 1e9, "unit": unit, "h2d_bytes_per_step": n * 4, "d2h_bytes_per_step": 4}
        cfg = {"workload": WORKLOADS["c2"],
               "kernel": ops.plan_info_reduce(ops.make_reduce_plan(elem, lv, n)).kernel.decode() + " (2 chained passes)"}

        def cpu():
            return cpu_baseline("c2", 4.0)
    elif w == "c4":
        Mm = Kk = 256
        batch = args.batch if args.batch != 1000000 else 8192
        t, tags, lv = c4_types()
        plan = ops.make_cgemul_plan(t, t, t, t, lv[-1], lv[-1], Mm, 1, Kk, tf_tags=tags, add_re=lv, form=TREE_VARIADIC)
        info = ops.plan_info_cgemul(plan)
        H = torch.empty(batch * Mm * Kk * 2, dtype=torch.int32, device=dev); xv = torch.empty(batch * Kk * 2, dtype=torch.int32, device=dev)
        ops.synth_fill(H, -(1 << 14), (1 << 14) - 1, 1, 0); ops.synth_fill(xv, -(1 << 14), (1 << 14) - 1, 1, 1)   # half range
        y = torch.empty(batch * Mm * 2, dtype=torch.int32, device=dev)
        ops.Qcgemul(y, H, xv, plan, batch=batch)
        want = oracle.cgemul_batched(plan, H[:Mm * Kk * 2].cpu().numpy(), xv[:Kk * 2].cpu().numpy(), 1)
        assert np.array_equal(y[:Mm * 2].cpu().numpy(), want), "complex mat-vec differs from the oracle"
        ms, launches, clocks = timed(lambda: ops.Qcgemul(y, H, xv, plan, batch=batch))
        units = batch * Mm * Kk
        value, unit, metric = world * units / (ms * 1e-3) / 1e9, "G complex-MAC/s", "bit-exact TFComplexMul (3-mult/5-add) complex mat-vec throughput, batched 256x256"
        bytes_alg = batch * (Mm * Kk + Kk + Mm) * 8
        roof = {"bound": "hbm", "achieved": bytes_alg / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "note": "near the HBM / integer-ALU balance point (8 B and ~25 integer ops per complex MAC)"}
        nb = min(batch, 512)
        hH = torch.empty(nb * Mm * Kk * 2, dtype=torch.int32).pin_memory(); hH.copy_(H[:nb * Mm * Kk * 2])
        hx = torch.empty(nb * Kk * 2, dtype=torch.int32).pin_memory(); hx.copy_(xv[:nb * Kk * 2])
        hy = torch.empty(nb * Mm * 2, dtype=torch.int32).pin_memory()
        e2e_s = wall(lambda: ops.Qcgemul(hy.numpy(), hH.numpy(), hx.numpy(), plan, batch=nb))
        e2e = {"value": nb * Mm * Kk / e2e_s / 1e9, "unit": unit, "h2d_bytes_per_step": nb * (Mm * Kk + Kk) * 8, "d2h_bytes_per_step": nb * Mm * 8}
        cfg = {"workload": WORKLOADS["c4"],
               "batch": batch, "kernel": info.kernel.decode()}

        def cpu():
            return cpu_baseline("c4", 8.0)
    else:  # c5
        Nn = Kk = 8192
        Mm = 8192 // world          # M-row panel of A and C per GPU; B is replicated.  A is handed over K-major (QgemulTransposedA:
                                    # stored K x M), so a row panel of op(A) is one contiguous block of the caller's matrix
        a = QType(7, 0); acc = QType(30, 0)
        plan = ops.make_gemul_plan(a, a, acc, Mm, Nn, Kk, mul=QType(15, 0), add=[acc], trans_a=True)
        info = ops.plan_info_gemul(plan)
        assert info.tensor_core == 1
        A = torch.empty(Mm * Kk, dtype=torch.int32, device=dev); B = torch.empty(Kk * Nn, dtype=torch.int32, device=dev)
        ops.synth_fill(A, -128, 127, 1, 0, rank * Mm * Kk); ops.synth_fill(B, -128, 127, 1, 1)
        Cd = torch.empty(Mm * Nn, dtype=torch.int32, device=dev)
        ops.Qgemul(Cd, A, B, plan)
        rows = torch.arange(0, Mm, 127, device=dev)                                  # exact float64 cross-check on sampled rows
        ref = (A.view(Mm, Kk)[rows].double() @ B.view(Nn, Kk).t().double())
        assert torch.equal(Cd.view(Nn, Mm).t()[rows].double(), ref), "tensor-core Qgemul differs from the exact float64 GEMM"
        ms, launches, clocks = timed(lambda: ops.Qgemul(Cd, A, B, plan))
        ops_total = 2.0 * Mm * Nn * Kk
        value, unit, metric = world * ops_total / (ms * 1e-3) / 1e12, "TOP/s", "bit-exact Qgemul 8192^3, int8-range operands, lossless accumulators (tcgen05 kind::i8), incl. int32->int8 pack"
        # library int8 GEMM on the same box = the measured tensor peak for this dtype
        A8 = torch.randint(-128, 128, (Mm, Kk), dtype=torch.int8, device=dev); B8 = torch.randint(-128, 128, (Kk, Nn), dtype=torch.int8, device=dev)
        best = 0.0
        try:
            for _ in range(3):
                torch._int_mm(A8, B8)
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); torch._int_mm(A8, B8); e1.record(); torch.cuda.synchronize()
                best = max(best, ops_total / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        except Exception:
            best = 0.0
        nominal = 2.0 * float(peaks.get("bf16_tflops", 1590.0))
        # the library GEMM under the SAME timing discipline as `value` (>= 1.2 s of back-to-back calls: both kernels run into the
        # power cap there), and the library GEMM fed from the reference's int32 words (torch narrowing passes + cuBLASLt)
        sus = sus_conv = 0.0
        try:
            ms_lib, _, clocks_lib = timed(lambda: torch._int_mm(A8, B8))
            sus = ops_total / (ms_lib * 1e-3) / 1e12
            ms_conv, _, _ = timed(lambda: torch._int_mm(A.view(Mm, Kk).to(torch.int8), B.view(Nn, Kk).to(torch.int8).t()))
            sus_conv = ops_total / (ms_conv * 1e-3) / 1e12
        except Exception:
            pass
        peak = sus if sus > 0 else (best if best > 0 else nominal)
        roof = {"bound": "tensor", "achieved": value / world, "peak": peak, "unit": "TOP/s", "scope": "per GPU",
                "peak_source": ("torch._int_mm (cuBLASLt int8, int8 operands) 8192^3 on this box, sustained: the same back-to-back timing loop as `value`" if sus > 0 else
                                "torch._int_mm (cuBLASLt int8) 8192^3 on this box, best single call" if best > 0 else "2 x measured bf16 (nominal i8:bf16 ratio)"),
                "peak_burst": best, "peak_burst_note": "best of 5 single cuBLASLt calls from idle (boost clocks, no power cap)",
                "library_from_int32_words": sus_conv, "library_from_int32_words_note": "torch int32 -> int8 narrowing of both operands + cuBLASLt, same loop: "
                "the like-for-like library figure for the reference's word format", "peak_2x_bf16": nominal}
        if sus > 0:
            roof["peak_clocks"] = clocks_lib
        # the same GEMM with the operands handed over as int8 words, K contiguous: the tensor maps sit on the caller's buffers,
        # no pack kernel runs (one launch per call)
        plan8 = ops.make_gemul_plan(a, a, acc, Mm, Nn, Kk, mul=QType(15, 0), add=[acc], trans_a=True, word_bytes=(1, 1, 0))
        A8w = A.to(torch.int8); B8w = B.to(torch.int8)
        ops.Qgemul(Cd, A8w, B8w, plan8)
        assert torch.equal(Cd.view(Nn, Mm).t()[rows].double(), ref), "zero-copy int8 Qgemul differs from the exact float64 GEMM"
        ms8, launches8, clocks8 = timed(lambda: ops.Qgemul(Cd, A8w, B8w, plan8))
        int8_line = {"value": world * ops_total / (ms8 * 1e-3) / 1e12, "unit": "TOP/s", "ms_per_step": ms8, "gpu_launches": launches8, "clocks": clocks8,
                     "note": "operands as int8 words, K contiguous (qb_qdesc.word_bytes = 1): read in place by TMA, no pack pass"}
        del A8w, B8w
        hA = torch.empty(Mm * Kk, dtype=torch.int32).pin_memory(); hA.copy_(A)
        hB = torch.empty(Kk * Nn, dtype=torch.int32).pin_memory(); hB.copy_(B)
        hC = torch.empty(Mm * Nn, dtype=torch.int32).pin_memory()
        e2e_s = wall(lambda: ops.Qgemul(hC.numpy(), hA.numpy(), hB.numpy(), plan))
        e2e = {"value": ops_total / e2e_s / 1e12, "unit": unit, "h2d_bytes_per_step": 2 * Mm * Kk * 4, "d2h_bytes_per_step": Mm * Nn * 4}
        cfg = {"workload": WORKLOADS["c5"],
               "kernel": info.kernel.decode(), "sharding": "M-row panels of A and C (%d rows per GPU, A stored K-major), B replicated, no collective" % Mm}

        def cpu():
            pl = ops.make_gemul_plan(a, a, acc, 16, 16, Kk, mul=QType(15, 0), add=[acc])
            Ah = np.ascontiguousarray(A.view(Mm, Kk)[:16].t().cpu().numpy()).reshape(-1); Bh = B[:16 * Kk].cpu().numpy()
            t0 = time.perf_counter(); oracle.gemul_batched(pl, Ah, Bh, 1); dt = time.perf_counter() - t0
            return {"value": 2.0 * 16 * 16 * Kk / dt / 1e12, "unit": unit, "cores": 1, "kind": "port", "sample": "a 16x16x8192 corner through the C oracle, 1 core"}

    roof["frac"] = roof["achieved"] / roof["peak"]
    if roof["bound"] == "hbm" and roof["frac"] > 1.0:
        roof["note"] = "the measured peak in MEASURED_PEAKS.json is a copy (read + write) bandwidth; this kernel only reads"
    roof["traffic"] = ncu_traffic(args.workload, bytes_alg) if args.workload in ("c2", "c4") else (ncu_traffic("c5", NCU_CAPTURE_ALG_BYTES["c5"]) if world == 1 else None)
    roof["traffic_source"] = "profiles/r1_%s_ncu_summary.json, scaled to this launch by algorithmic bytes" % args.workload
    roof.setdefault("peak_source", "MEASURED_PEAKS.json" if peaks else "fallback")
    if rank != 0:
        dist.destroy_process_group()
        return
    if w in ("c2", "c4", "c5") and os.path.exists(os.path.join(ROOT, "profiles", "r2_%s_ncu_summary.json" % w)):
        roof["traffic_source"] = roof["traffic_source"].replace("r1_", "r2_")
    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong" if w == "c5" else "weak", "vs_baseline": None,
            "dtype": "int32" if w != "c5" else "int8 x int8 -> int32",
            "data": "synthetic", "config": cfg, "roofline": roof, "clocks": clocks, "e2e": e2e, "gpu_launches": launches}
    if w == "c5":
        line["int8_operands"] = int8_line
    if not args.no_cpu and world == 1:
        try:
            line["cpu_baseline"] = cpu()
        except Exception as ex:      # e.g. oracle/_ref absent
            line["cpu_baseline"] = {"unavailable": str(ex)}
    if embed:
        return line
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=1000000, help="instances per GPU (BASELINE config 3: 1e6)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-chunk", type=int, default=65536, help="instances per host-buffer API call")
    ap.add_argument("--no-ref-words", action="store_true", help="skip the supplementary end-to-end measurement with the reference's int32 words")
    ap.add_argument("--no-sweep", action="store_true", help="skip the supplementary device-resident Monte-Carlo sweep")
    ap.add_argument("--e2e-instances", type=int, default=1000000, help="instances per end-to-end step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-secondary", action="store_true", help="skip the BASELINE config 2 / 4 / 5 legs appended to the default single-GPU line")
    ap.add_argument("--check-instances", type=int, default=512, help="instances per rank compared with the oracle before timing")
    ap.add_argument("--workload", default="c3", choices=["c3", "c2", "c4", "c5"],
                    help="c3 = headline (batched 64^3 Qgemul); c2 = Qreduce 2^24; c4 = TFComplexMul 256x256 mat-vec; c5 = 8192^3 tcgen05 i8")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c3":
        run_ours(args)
    else:
        run_secondary(args)


if __name__ == "__main__":
    main()
