#!/usr/bin/env python
"""Instruction / stall-sample share per pass of k_tile_fast (sections are found by their comment markers).
usage: pass_profile.py <report.ncu-rep> <cubin> [template-arg e.g. ILb1EE]"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rep, cubin = sys.argv[1:3]
inst_tag = sys.argv[3] if len(sys.argv) > 3 else ""
kname = "k_tile_fast"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
blk = [b for b in blocks if kname in b["name"]][0]
hdr = blk["rows"][0]
si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
insts = [(int(r[si] or 0), int(r[ii] or 0)) for r in blk["rows"][1:] if len(r) > ii]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
secs = re.split(r"\n(?=\s*\.section\s+\.text\.)", dis)
seq = None
for sec in secs:
    head = sec.split("\n")[0]
    if kname not in head or inst_tag not in head:
        continue
    s, line = [], None
    for ln in sec.splitlines():
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            line = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln):
            s.append(line)
    if len(s) == len(insts):
        seq = s
        break
assert seq is not None, "no matching instantiation"
src = open("/root/repo/vaporetto_b200/csrc/kernels.cu").read().splitlines()


def find(t):
    return next(i + 1 for i, l in enumerate(src) if t in l)


marks = [("setup", find("CTA-shared tables ---")), ("tables", find("per-sentence tables ---")),
         ("range", find("choose the longest sentence range")), ("stage", find("stage the range's bytes")),
         ("passA", find("pass A:")), ("passB", find("pass B:")), ("passC", find("pass C:")), ("passD", find("pass D:")),
         ("end", find("// k_score_general"))]
first_kernel_line = find("k_tile_fast — persistent")
agg = defaultdict(lambda: [0, 0])
for (smp, ins), ln in zip(insts, seq):
    if ln is None:
        key = "?"
    elif ln[0] != "kernels.cu":
        key = "inlined:" + ln[0]
    else:
        L = ln[1]
        key = "inlined helpers (probe/gather/decode/scan)"
        for (nm, st), (_, en) in zip(marks, marks[1:]):
            if st <= L < en:
                key = nm
    agg[key][0] += smp
    agg[key][1] += ins
tot = sum(a[0] for a in agg.values())
toti = sum(a[1] for a in agg.values())
print(f"{blk['name']}: {toti} warp instructions, {tot} samples")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:44s} inst {100 * a[1] / toti:5.1f}%  samples {100 * a[0] / tot:5.1f}%")
