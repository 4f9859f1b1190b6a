#!/usr/bin/env python
"""SASS / resource evidence of the shipped library, produced WITHOUT a GPU (cuobjdump works on the cross-compiled .so):

    python scripts/sass_report.py profiles/r02_sass_mnemonics.txt profiles/r02_kernel_resources.txt

File 1: counts of the Blackwell-specific mnemonics (UTMALDG / UTMASTG = cp.async.bulk.tensor load / store, UTCHMMA =
tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, ...) over the whole library and per tensor-core kernel.
File 2: registers / shared memory / spill bytes per kernel from `cuobjdump -res-usage`.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "efficientat_b200", "libeat_b200.so")
MNEMONICS = ["UTMALDG", "UTMASTG", "UTMAPF", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKPF", "UTCATOMSWS", "SYNCS", "FFMA2",
             "LDGSTS", "ATOMS"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main(sass_out, res_out):
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    total = collections.Counter()
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            op = m.group(1)
            for k in MNEMONICS:
                if op == k or op.startswith(k + "."):
                    total[k] += 1
                    per[cur][k] += 1
    names = demangle(list(per))
    with open(sass_out, "w") as f:
        for k in MNEMONICS:
            f.write(f"{k:12s} {total[k]}\n")
        f.write(f"\nkernels in the library: {len(per)}\n")
        f.write("\nper kernel with tensor-core / TMA instructions (UTMALDG / UTMASTG / UTCHMMA / LDTM):\n")
        for fn, c in per.items():
            if c["UTMALDG"] or c["UTMASTG"] or c["UTCHMMA"] or c["LDTM"]:
                f.write(f"{c['UTMALDG']:6d} {c['UTMASTG']:6d} {c['UTCHMMA']:6d} {c['LDTM']:6d}  {names[fn][:150]}\n")
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    rows = []
    fn = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and fn:
            rows.append((fn, int(m.group(1)), int(m.group(2)), int(m.group(3)), int(m.group(4))))
            fn = None
    names = demangle([r[0] for r in rows])
    with open(res_out, "w") as f:
        f.write("registers  stack  static-smem  local   kernel   (cuobjdump -res-usage; dynamic shared memory is set at launch)\n")
        for fn, reg, stack, sh, loc in sorted(rows, key=lambda r: names[r[0]]):
            f.write(f"{reg:9d} {stack:6d} {sh:12d} {loc:6d}   {names[fn][:170]}\n")
    print(f"{len(per)} kernels; UTMALDG {total['UTMALDG']} UTMASTG {total['UTMASTG']} UTCHMMA {total['UTCHMMA']} LDTM {total['LDTM']}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
