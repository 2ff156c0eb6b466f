#!/usr/bin/env python3
"""SASS listings of the headline kernels for profiles/ (evidence a reader can grep without rebuilding):
   python tools/sass_listing.py            -> profiles/r2_sass_<kernel>.txt + profiles/r2_sass_index.json
Each listing is `cuobjdump -sass` of one kernel of qublas_b200/libqb200.so (addresses and encodings stripped), preceded by
its opcode histogram and resource usage."""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "qublas_b200", "libqb200.so")
KERNELS = {                               # tag -> substring of the mangled name
    "c3_k_gemul_tile": "k_gemul_tileI5CfgC3",
    "c5_k_i8gemm_pair": "k_i8gemm_pairE",
    "c5_k_i8gemm_pair_fused": "k_i8gemm_pair_fusedE",
    "c2_k_reduce_ct_pass1": "k_reduce_ctI7CfgC2b0Li0ELi12E",
    "c4_k_cgemv_rows": "k_cgemv_rowsI5CfgC4",
}
PROOF = ["UTCIMMA", "UTMALDG", "LDTM", "UTCBAR", "UBLKCP", "SYNCS", "FFMA", "IADD3", "LEA.HI", "SHF", "LDS.128", "STL", "LDL", "LDG.E.128", "VIMNMX", "VIMNMX3", "I2FP"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    funcs, cur = {}, None
    for l in txt.splitlines():
        m = re.search(r"Function : (\S+)", l)
        if m:
            cur = m.group(1); funcs[cur] = []
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*/\*", l)
        if m:
            funcs[cur].append(m.group(2).rstrip(" ;") )
    usage = {}
    rl = res.splitlines()
    for i, l in enumerate(rl):
        m = re.search(r"Function (\S+):", l)
        if m and i + 1 < len(rl):
            usage[m.group(1)] = rl[i + 1].strip()
    index = {}
    for tag, pat in KERNELS.items():
        names = [n for n in funcs if pat in n]
        if not names:
            print("missing", tag, file=sys.stderr)
            continue
        name = names[0]
        ins = funcs[name]
        hist = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", x).split()[0] for x in ins if x)
        path = os.path.join(ROOT, "profiles", "r2_sass_%s.txt" % tag)
        with open(path, "w") as f:
            f.write("# %s\n# %s\n# %d instructions; opcode histogram:\n" % (name, usage.get(name, ""), len(ins)))
            for k, v in hist.most_common():
                f.write("#   %6d  %s\n" % (v, k))
            f.write("\n".join(ins) + "\n")
        index[tag] = {"function": name, "resources": usage.get(name, ""), "instructions": len(ins),
                      "proof": {k: sum(v for op, v in hist.items() if op.startswith(k)) for k in PROOF if any(op.startswith(k) for op in hist)}}
        print(tag, len(ins), "instructions ->", os.path.relpath(path, ROOT))
    json.dump(index, open(os.path.join(ROOT, "profiles", "r2_sass_index.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
