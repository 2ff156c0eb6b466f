"""Instruction mix of one kernel in a cubin / shared library: python tools/sass_mix.py <file> <substring of the mangled name>
Prints the opcode histogram (whole kernel) -- read next to the ncu instruction counts, which weight by execution."""
import collections
import re
import subprocess
import sys


def main():
    path, pat = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    fn, c = None, collections.Counter()
    for l in txt.splitlines():
        m = re.search(r"Function : (\S+)", l)
        if m:
            fn = m.group(1)
            continue
        if fn is None or pat not in fn:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
        if m:
            c[m.group(1)] += 1
    print(sum(c.values()), "instructions")
    for k, v in c.most_common(40):
        print(f"{v:7d}  {k}")


if __name__ == "__main__":
    main()
