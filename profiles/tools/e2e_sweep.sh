#!/bin/bash
# e2e throughput of vpt_predict_batch vs pipeline chunk size (run on the GPU box)
for c in 16384 32768 65536 131072 262144; do
  VPT_CHUNK_SENTENCES=$c python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunk', $c, 'e2e MB/s', d['e2e']['value'])"
done
