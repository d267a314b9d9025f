#!/usr/bin/env python
"""Instruction histogram per kernel instantiation of frave_b200/libfri_cuda.so, from `cuobjdump -sass`,
plus registers / spills / shared memory from `cuobjdump -res-usage`.  Makes the TMA (UBLKCP), cp.async
(LDGSTS), byte-gather (LDS.U8 / STS.U8) and no-spill claims checkable without rebuilding.

    python profiles/sass_histogram.py [lib.so] > profiles/r2_sass_histogram.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "frave_b200", "libfri_cuda.so")
WATCH = ["UBLKCP", "UBLKPF", "LDGSTS", "LDG", "STG", "LDS", "STS", "LDS.U8", "STS.U8", "LDS.128", "STS.128", "LDG.E.128",
         "STG.E.128", "SHFL", "VIMNMX", "PRMT", "IMAD", "IADD3", "LOP3", "SHF", "LEA", "BAR", "SYNCS", "ATOMS", "RED", "CCTL",
         "LDL", "STL", "BRA", "ACQBULK", "UTMASTG", "UTMALDG"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|fri::|unnamed>::", "", name)
    name = re.sub(r"\(.*", "", name)
    return name.replace("unsigned char", "u8").replace("unsigned short", "u16").replace("(bool)", "")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)", line)
        if m and cur:
            usage[cur] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
    funcs = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            op = m.group(1)
            c = funcs[cur]
            c["_total"] += 1
            c[op.split(".")[0]] += 1
            if op.startswith(("LDS.U8", "STS.U8", "LDS.128", "STS.128", "LDG.E.128", "STG.E.128")):
                c[".".join(op.split(".")[:3]) if op[1:3] == "DG" or op[1:3] == "TG" else ".".join(op.split(".")[:2])] += 1
    names = demangle(list(funcs))
    print(f"# {os.path.relpath(LIB, ROOT)}: static SASS instruction counts per kernel (cuobjdump -sass), "
          f"registers / static shared / local (spill) bytes (cuobjdump -res-usage)")
    print("# columns: total | regs smem local | " + " ".join(WATCH))
    for f, c in funcs.items():
        r = usage.get(f, (None, None, None))
        print(f"{short(names[f])}")
        print(f"    total={c['_total']} regs={r[0]} static_smem={r[1]} local={r[2]} | " +
              " ".join(f"{k}={c[k]}" for k in WATCH if c[k]))


if __name__ == "__main__":
    main()
