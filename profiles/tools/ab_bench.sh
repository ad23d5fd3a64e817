#!/bin/bash
# usage (GPU box): profiles/tools/ab_bench.sh <rounds> <lib.so> [<lib.so> ...]   ("-" = the in-tree library)
# Device-timed ms/step of bench.py for several builds of the library, interleaved on one box.
rounds=$1; shift
for r in $(seq 1 $rounds); do
  for lib in "$@"; do
    if [ "$lib" = "-" ]; then unset VPT_B200_LIBRARY; else export VPT_B200_LIBRARY=$PWD/$lib; fi
    ms=$(python bench.py --no-cpu-baseline --e2e-steps 1 ${AB_ARGS} 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], 'single_call_us', d['e2e'].get('single_call_us'))")
    echo "round $r  $lib  ms_per_step=$ms"
  done
done
