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
    # the lines path: splitter, trims through every scoring kernel, look-back output kernel; tiny and default chunks
    lines = [text[int(offs[i]):int(offs[i + 1])].tobytes() for i in range(len(lens))]
    data = b"\r\n".join(lines[:7]) + b"\n\n\xff\xfe\na/b c\\d.\n" + b"\n".join(lines[7:])
    for chunk in ("64", "4096", ""):
        os.environ["VPT_CHUNK_BYTES"] = chunk
        for no_norm in (True, False):
            got, nl = p.tokenize_lines(data, no_norm=no_norm)
            want, wl = o.tokenize_lines(data, no_norm=no_norm)
            assert nl == wl and got.tobytes() == want, (kw, chunk, no_norm)
    os.environ["VPT_CHUNK_BYTES"] = "4096"
    got, nl = p.tokenize_lines(data, wsconst="GD")          # k_grapheme + k_wsconst
    want, wl = o.tokenize_lines(data, wsconst="GD")
    assert nl == wl and got.tobytes() == want, kw
    cr = p.predict_batch_compact(text, offs)                # k_pack_bits, k_sentence_info, k_token_scan, k_token_base
    assert np.array_equal(cr.boundaries(), r.boundaries), kw
    if tags:
        res, tok, cand, uns = p.predict_batch_tags(text, offs)      # k_tags (single kernel)
        ct = p.predict_batch_compact(text, offs, tags=True)         # k_tags<locate> + k_tok_lookup + k_tok_score
        assert np.array_equal(ct.boundaries(), res.boundaries) and ct.token_ids.size == int(ct.n_tokens.sum())
        assert int((ct.token_ids >= 0).sum()) == int((tok >= 0).sum())
        for no_norm in (True, False):                               # k_tok_write_tags
            got, nl = p.tokenize_lines(data, no_norm=no_norm, predict_tags=True)
            want, wl = o.tokenize_lines(data, no_norm=no_norm, predict_tags=True)
            assert nl == wl and got.tobytes() == want, (kw, no_norm)
    one = vb.Sentence.from_raw("まぁ社長は火星猫だ")                 # vpt_predict: zero-copy single-sentence path
    p.predict(one)
print("sanitizer workload ok")
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 60 python /tmp/san_small.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitizer workload ok|Error|hazard" gpurun_out/sanitize_$tool.log | head -12
done
