#!/usr/bin/env python
"""Static SASS statistics of one kernel of an object file / shared library, without a GPU:
instructions per source line (needs -lineinfo), per opcode class, and per pipe.

usage: sass_lines.py <obj-or-so> <kernel-substr> [file-substr [line_lo line_hi]]
Prints: total instructions, instructions whose source line falls into [line_lo, line_hi] of <file-substr>
grouped by line, and the opcode histogram of that region.  The pipe split (alu / fma / lsu / other) uses the
B300_MICROARCH.md table: IMAD/FFMA on the fma pipe, IADD3/LOP3/SHF/PRMT/ISETP/SEL on the alu pipe.
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

FMA = ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "IMUL")
LSU = ("LDG", "STG", "LDS", "STS", "LDL", "STL", "ATOM", "RED", "LDSM", "LD.", "ST.")
XU = ("POPC", "FLO", "BREV", "MUFU")


def pipe(op):
    base = op.split(".")[0]
    if base in FMA:
        return "fma"
    if any(op.startswith(x) for x in LSU) or base in ("LD", "ST"):
        return "lsu"
    if base in XU:
        return "xu"
    if base in ("SHFL", "VOTE", "VOTEU", "MATCH", "REDUX"):
        return "shfl/vote"
    if base in ("BRA", "BSSY", "BSYNC", "EXIT", "WARPSYNC", "BAR", "NANOSLEEP", "CALL", "RET", "BREAK", "YIELD", "NOP"):
        return "ctrl"
    return "alu"


def main():
    obj, kname = sys.argv[1], sys.argv[2]
    fsub = sys.argv[3] if len(sys.argv) > 3 else None
    lo = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    hi = int(sys.argv[5]) if len(sys.argv) > 5 else 10 ** 9
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, check=True, capture_output=True)
        cubins = [os.path.join(td, f) for f in os.listdir(td) if f.endswith(".cubin")]
        dis = ""
        for cb in cubins:
            dis += subprocess.run(["nvdisasm", "-g", "-c", cb], capture_output=True, text=True).stdout
    secs = re.split(r"\n(?=\s*\.section\s+\.text\.)", dis)
    sec = [s for s in secs if kname in s.split("\n")[0]]
    if not sec:
        sys.exit("kernel not found")
    sec = sec[0]
    cur = None
    per_line = collections.Counter()
    ops = collections.Counter()
    pipes = collections.Counter()
    total = 0
    for ln in sec.splitlines():
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if not m:
            continue
        total += 1
        if cur and (fsub is None or fsub in cur[0]) and lo <= cur[1] <= hi:
            per_line[cur] += 1
            ops[m.group(2).split(".")[0]] += 1
            pipes[pipe(m.group(2))] += 1
    print(f"kernel section: {sec.splitlines()[0].strip()[:120]}")
    print(f"total static instructions: {total}; in region: {sum(per_line.values())}")
    for (f, l), n in sorted(per_line.items()):
        print(f"  {f}:{l}  {n}")
    print("opcodes:", ", ".join(f"{k} {v}" for k, v in ops.most_common()))
    print("pipes:", dict(pipes))


if __name__ == "__main__":
    main()
