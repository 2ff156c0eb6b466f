#!/usr/bin/env python3
"""Condense an .ncu-rep (ncu --set full) into the handful of numbers DESIGN.md / bench.py cite.
usage: python tools/ncu_summary.py gpurun_out/r1_c3.ncu-rep [units_per_launch] > profiles/r1_c3_summary.json"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_hw_peak",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "pipe_tensor_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__cycles_elapsed.max": "sm_cycles",
}


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else None
    txt = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, unit = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k, name in KEYS.items():
            if k in hdr:
                i = hdr.index(k)
                try:
                    d[name] = float(r[i])
                except ValueError:
                    d[name] = r[i]
                d[name + "_unit"] = unit[i]
        if "dram_read" in d and "dram_write" in d:
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            d["traffic_bytes"] = d["dram_read"] * scale.get(d["dram_read_unit"], 1) + d["dram_write"] * scale.get(d["dram_write_unit"], 1)
        if units and "warp_instructions" in d:
            d["thread_instructions_per_unit"] = d["warp_instructions"] * 32 / units
        out.append(d)
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
