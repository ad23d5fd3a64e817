#!/bin/bash
# usage (on the GPU box, via gpurun): profiles/run_ncu_lines.sh <kernel-regex> <out-name> [chunk MiB]
# One `ncu --set full` capture of one launch of a kernel of the vpt_tokenize_lines path (tools/lines_sweep.py).
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$1 -s 3 -c 1 -o gpurun_out/$2 -f \
    python tools/lines_sweep.py 1000000 ${3:-128} > gpurun_out/$2.log 2>&1
ls -la gpurun_out/$2.ncu-rep
