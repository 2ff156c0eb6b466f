#!/usr/bin/env python3
"""Per-opcode stall breakdown of one kernel from an .ncu-rep captured with --import-source on.
usage: python tools/ncu_stalls.py rep.ncu-rep [macs] > profiles/xx_stalls_by_opcode.json"""
import collections
import csv
import io
import json
import re
import subprocess
import sys


def main():
    rep = sys.argv[1]
    macs = float(sys.argv[2]) if len(sys.argv) > 2 else None
    txt = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(io.StringIO(txt)))
    hi = [i for i, r in enumerate(rows) if "Source" in r][0]
    h = rows[hi]
    stall_cols = [(i, k[6:]) for i, k in enumerate(h) if k.startswith("stall_") and "Not Issued" not in k]
    tot = collections.Counter(); by_op = collections.defaultdict(collections.Counter); inst = collections.Counter(); samples = collections.Counter()
    wf = collections.Counter(); wf_ideal = collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) < len(h):
            continue
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[h.index("Source")])
        if not m:
            continue
        op = m.group(1)
        inst[op] += int(r[h.index("Instructions Executed")] or 0)
        samples[op] += int(r[h.index("# Samples")] or 0)
        wf[op] += int(r[h.index("L1 Wavefronts Shared")] or 0); wf_ideal[op] += int(r[h.index("L1 Wavefronts Shared Ideal")] or 0)
        for i, k in stall_cols:
            v = int(r[i] or 0)
            tot[k] += v; by_op[op][k] += v
    n = sum(tot.values())
    out = {"kernel": rows[0][1], "warp_samples": n, "warp_instructions": sum(inst.values()),
           "stall_share": {k: round(v / n, 4) for k, v in sorted(tot.items()) if v},
           "smem_wavefronts": {k: {"wavefronts": v, "ideal": wf_ideal[k]} for k, v in wf.items() if v}}
    if macs:
        out["per_mac_warp_instructions"] = {k: round(v * 32 / macs, 4) for k, v in inst.most_common(16)}
    out["by_opcode"] = {k: {"samples": samples[k], "top_stalls": dict(by_op[k].most_common(4))} for k, _ in samples.most_common(8)}
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
