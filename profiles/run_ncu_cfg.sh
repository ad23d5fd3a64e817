#!/bin/bash
# usage: profiles/run_ncu_cfg.sh <config> <kernel-regex> <out-name>
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$2 -s 1 -c 1 -o gpurun_out/$3 -f \
    python bench.py --config $1 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/$3.log 2>&1
ls -la gpurun_out/$3.ncu-rep
