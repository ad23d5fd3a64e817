#!/bin/bash
# usage (GPU box): profiles/tools/ab_tags.sh <rounds> <lib.so> [<lib.so> ...]   ("-" = the in-tree library)
# Wall time of vpt_predict_batch_compact with tags (config 3, 1 M sentences, third call) for several builds, interleaved.
rounds=$1; shift
for r in $(seq 1 $rounds); do
  for lib in "$@"; do
    if [ "$lib" = "-" ]; then unset VPT_B200_LIBRARY; else export VPT_B200_LIBRARY=$PWD/$lib; fi
    ms=$(VPT_TRACE=0 python profiles/tools/trace_compact.py 1000000 tags 2>&1 | grep "compact call 2" | sed 's/.*: //')
    echo "round $r  $lib  compact+tags call: $ms"
  done
done
