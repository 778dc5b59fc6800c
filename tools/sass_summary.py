#!/usr/bin/env python
"""Per-kernel opcode counts of the shipped library (cuobjdump -sass): the evidence that the step kernels are sm_100a
code using the TMA engine (UTMALDG = cp.async.bulk.tensor / gather4, UBLKCP = cp.async.bulk), mbarriers (SYNCS),
cp.async (LDGSTS) and 128-bit accesses.  Writes a table to stdout:

    python tools/sass_summary.py [finenvs_b200/libfinenvs_b200.so] > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "finenvs_b200", "libfinenvs_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()  # noqa: E731

COLS = ["UTMALDG", "UBLKCP.S.G", "UBLKCP.G.S", "SYNCS", "LDGSTS", "STS", "LDS", "LDG.E.128", "STG.E.128", "FENCE.VIEW.ASYNC",
        "ATOMS", "SHFL", "DADD/DMUL", "total"]
archs, kernels, cur = set(), collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        archs.add(m.group(1))
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m or cur is None:
        continue
    op = m.group(1)
    cur["total"] += 1
    for c in COLS[:-2]:
        if op.startswith(c):
            cur[c] += 1
    if op.startswith("UBLKCP"):
        cur["UBLKCP.S.G" if ".S.G" in line else "UBLKCP.G.S"] += 0  # counted above by prefix when the suffix matches
    if op.startswith(("DADD", "DMUL")):
        cur["DADD/DMUL"] += 1

print(f"{os.path.relpath(lib, ROOT)}: arch = {', '.join(sorted(archs))}; instructions per kernel (static SASS counts)\n")
print(f"{'kernel':<72}" + "".join(f"{c:>17}" for c in COLS))
tot = collections.Counter()
for name, c in kernels.items():
    short = re.sub(r"\(anonymous namespace\)::", "", demangle(name))
    short = re.sub(r"\(.*", "", short).replace("void ", "")
    print(f"{short[:71]:<72}" + "".join(f"{c[col]:>17}" for col in COLS))
    tot.update(c)
print(f"{'ALL KERNELS':<72}" + "".join(f"{tot[col]:>17}" for col in COLS))
