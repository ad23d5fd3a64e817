#!/usr/bin/env python
"""Maps the per-SASS-instruction samples of an `ncu --import-source on` report to CUDA source lines.

usage: line_profile.py <report.ncu-rep> <kernel-name-substring> <cubin from `cuobjdump -xelf all lib.so`> [top]
Needs ncu and nvdisasm on PATH (no GPU).  Instructions are matched by order of appearance.
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rep, kname, cubin = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# find kernel block
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
ti = hdr.index("Thread Instructions Executed")
insts = [(r[1].strip(), int(r[si] or 0), int(r[ii] or 0), int(r[ti] or 0)) for r in blk["rows"][1:] if len(r) > ii]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
seq = None
for sec in re.split(r"\n(?=\s*\.section\s+\.text\.)", dis):
    if kname not in sec.split("\n")[0]:
        continue
    cand, line = [], None
    for ln in sec.splitlines():
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            line = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln):
            cand.append(line)
    if seq is None or len(cand) == len(insts):
        seq = cand
    if len(cand) == len(insts):
        break
if len(seq) != len(insts):
    print(f"warning: {len(seq)} disassembled vs {len(insts)} profiled instructions", file=sys.stderr)
agg = defaultdict(lambda: [0, 0, 0])
for (sass, smp, ins, tins), ln in zip(insts, seq):
    a = agg[ln]
    a[0] += smp
    a[1] += ins
    a[2] += tins
tot = sum(a[0] for a in agg.values()) or 1
toti = sum(a[1] for a in agg.values()) or 1
src = {}
print(f"kernel {blk['name']}: {tot} samples, {toti} warp instructions")
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][int(__import__('os').environ.get('SORT_INST','0'))])[:top]:
    text = ""
    if ln:
        try:
            if ln[0] not in src:
                import glob
                cands = glob.glob(f"/root/repo/vaporetto_b200/csrc/{ln[0]}")
                src[ln[0]] = open(cands[0]).read().splitlines() if cands else []
            text = src[ln[0]][ln[1] - 1].strip()[:90]
        except Exception:
            pass
    print(f"{100 * a[0] / tot:5.1f}% smp {100 * a[1] / toti:5.1f}% inst  avg_thr={a[2] / max(a[1], 1):4.1f}  {ln}  {text}")
