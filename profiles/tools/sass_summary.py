#!/usr/bin/env python
"""Static summary of the kernels inside the built library (no GPU needed): architecture of the embedded cubins, and per kernel
the resource usage and the counts of the instructions that show how it moves data (UBLKCP = TMA bulk copy, SYNCS = mbarrier,
LDG...ENL2.256 = 32-byte record loads that bypass L1, STG.E.EF = streaming stores, SHFL / VOTE / BAR / ATOMS / ATOMG / RED).
usage: sass_summary.py [library.so]  > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "..", "vaporetto_b200", "libvaporetto_b200.so")
elf = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
print("embedded cubins:", ", ".join(sorted(set(re.findall(r"sm_\w+", elf)))))
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
for m in re.finditer(r"Function (\S+):\n\s*(.*)", res):
    usage[m.group(1)] = m.group(2)
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
name = None
counts = collections.defaultdict(collections.Counter)
for ln in sass.splitlines():
    m = re.match(r"\s+Function : (\S+)", ln)
    if m:
        name = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and name:
        op = m.group(1)
        c = counts[name]
        c["instructions"] += 1
        base = op.split(".")[0]
        if base in ("UBLKCP", "SYNCS", "SHFL", "VOTE", "BAR", "ATOMS", "ATOMG", "RED", "LDS", "STS", "LDL", "STL", "NANOSLEEP"):
            c[base] += 1
        if base == "LDG":
            c["LDG"] += 1
            if ".256" in op:
                c["LDG.256" + (" (ENL2: L1 no-allocate)" if "ENL2" in op else "")] += 1
        if base == "STG":
            c["STG"] += 1
            if ".EF" in op:
                c["STG.EF (streaming)"] += 1
demangle = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines() if counts else []
for mangled, pretty in zip(counts, demangle):
    short = re.sub(r"vpt::\(anonymous namespace\)::|\(vpt::[A-Za-z]+(, vpt::[A-Za-z]+)*\)|void ", "", pretty)
    c = counts[mangled]
    print(f"\n{short}\n  {usage.get(mangled, '').strip()}")
    print("  " + ", ".join(f"{k} {v}" for k, v in c.items()))
