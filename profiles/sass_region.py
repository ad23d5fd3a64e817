#!/usr/bin/env python
"""Dumps the SASS of a source-line range of a kernel with per-instruction execution counts.
usage: sass_region.py <report.ncu-rep> <kernel-substr> <cubin> <file> <line_lo> <line_hi>"""
import csv, re, subprocess, sys
rep, kname, cubin, fname, lo, hi = sys.argv[1:7]
lo, hi = int(lo), int(hi)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
blk = [b for b in blocks if kname in b["name"]][0]
hdr = blk["rows"][0]
si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
insts = [(r[1].strip(), int(r[si] or 0), int(r[ii] or 0)) for r in blk["rows"][1:] if len(r) > ii]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
seq = None
for sec in re.split(r"\n(?=\s*\.section\s+\.text\.)", dis):
    if kname not in sec.split("\n")[0]:
        continue
    cand, line = [], None
    for ln in sec.splitlines():
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            line = (m.group(1).split("/")[-1], int(m.group(2))); continue
        if re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln):
            cand.append(line)
    if len(cand) == len(insts):
        seq = cand; break
tot = 0
for (sass, smp, ins), ln in zip(insts, seq):
    if ln and ln[0] == fname and lo <= ln[1] <= hi:
        print(f"{ln[1]:5d} {ins:10d} {smp:6d}  {sass[:100]}")
        tot += ins
print("total inst in region:", tot)
