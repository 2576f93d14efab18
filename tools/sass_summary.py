#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of the shipped library (cuobjdump -sass):
the evidence that the hot kernels are sm_100a code using the FP64 tensor pipe
(DMMA), TMA bulk copies (UBLKCP), mbarriers (SYNCS), 256-bit global moves
(LDG/STG .256), cp.async (LDGSTS) and distributed shared memory stores.

    python tools/sass_summary.py [lib.so] > profiles/r2_sass_summary.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "bounded_lsq_b200", "libblsq_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
PAT = collections.OrderedDict([
    ("DMMA", r"\bDMMA"), ("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"),
    ("LDG.256", r"\bLDG\S*\.256"), ("STG.256", r"\bSTG\S*\.256"),
    ("LDGSTS", r"\bLDGSTS"), ("DSMEM st", r"\bST\.E\S*\s|\bSTAS"), ("UCGABAR", r"\bUCGABAR|\bCGABAR"),
    ("DFMA", r"\bDFMA"), ("MUFU", r"\bMUFU"), ("SHFL", r"\bSHFL"), ("LDL+STL", r"\b(LDL|STL)\b"),
])
rows = []
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"^void ", "", name)
        name = name[:name.find("(")] if "(" in name else name
        cur = [name, 0, collections.Counter()]
        rows.append(cur)
        continue
    if cur is not None and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        cur[1] += 1
        for k, p in PAT.items():
            if re.search(p, line):
                cur[2][k] += 1
print("# SASS summary of %s (sm_100a, cuobjdump -sass)\n" % os.path.basename(lib))
tot = collections.Counter()
for _, _, c in rows:
    tot.update(c)
print("Totals: " + ", ".join("%s %d" % (k, tot[k]) for k in PAT) + "\n")
print("| kernel | instructions | " + " | ".join(PAT) + " |")
print("|---|---|" + "---|" * len(PAT))
for name, n, c in sorted(rows, key=lambda r: r[0]):
    print("| `%s` | %d | " % (name, n) + " | ".join(str(c[k]) if c[k] else "" for k in PAT) + " |")
