#!/bin/bash
# usage (on the GPU box, via gpurun): profiles/run_ncu.sh <kernel-regex> <out-name> [bench args]
# One `ncu --set full` capture of the 2nd launch of the kernel during a short bench run.  The library that was
# profiled is kept beside the report (gpurun_out/<out-name>.so) so that the SASS can be mapped to source lines later.
mkdir -p gpurun_out
cp vaporetto_b200/libvaporetto_b200.so gpurun_out/$2.so
ncu --set full --clock-control none --import-source on -k regex:$1 -s 1 -c 1 -o gpurun_out/$2 -f \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 ${@:3} > gpurun_out/$2.log 2>&1
ls -la gpurun_out/$2.ncu-rep
