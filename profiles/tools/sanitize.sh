#!/bin/bash
# compute-sanitizer memcheck + racecheck over a small parity workload (run on the GPU box)
mkdir -p gpurun_out
cat > /tmp/san_small.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import vaporetto_b200 as vb
from vpt_testlib import synth
from vpt_testlib.oracle import OraclePredictor
for kw, tags in ((dict(), False), (dict(dict_words=2000), False), (dict(tag_models=100), True)):
    mb = synth.gen_model_bccwj_shaped(n_patterns=4000, sample_sentences=8000, **kw)
    p, o = vb.Predictor(vb.Model.read(mb), predict_tags=tags), OraclePredictor(mb, predict_tags=tags)
    lens = [1, 2, 40, 3100, 5, 700] + [37] * 300
    cps = synth.gen_codepoints(len(lens), np.array(lens), seed=5)
    text, offs = synth.encode_utf8(cps, np.array(lens))
    r = p.predict_batch(text, offs, want_states=tags)
    sc, bd, boff, st = o.predict_batch(text, offs, nthreads=2)
    assert np.array_equal(r.scores, sc) and np.array_equal(r.boundaries, bd), kw
print("sanitizer workload ok")
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_small.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitizer workload ok|Error|hazard" gpurun_out/sanitize_$tool.log | head -12
done
