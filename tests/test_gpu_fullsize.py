"""GPU parity at the sizes BASELINE.json states (configs 2, 3, 4): 300 000 char patterns; + 20 000 tag models with the
pattern-id states; + a 500 000-word dictionary -- over >= 100 000 synthetic sentences each, compared with the CPU
oracle: scores, boundaries, offsets, status and (config 3) both state arrays, every element.  Plus `fill_tags` strings
through the GPU states on a synthetic tag model against the oracle's predict_tags (reference predictor.rs:546-637,
known-answer shape :863-903).  Bit-exact: all arithmetic is i32 / u32.
"""
import os

import numpy as np
import pytest

import vaporetto_b200 as vb
from vpt_testlib import synth
from vpt_testlib.oracle import OraclePredictor

pytestmark = pytest.mark.gpu

N_SENT = 120_000
NTHREADS = max(1, min(64, os.cpu_count() or 1))


def _text(seed):
    return synth.gen_text(N_SENT, 40, seed=synth.TEXT_SEED + seed)


def _compare_batch(p, o, text, offs, states=False):
    r = p.predict_batch(text, offs, want_states=states)
    sc, bd, boff, st = o.predict_batch(text, offs, nthreads=NTHREADS)
    assert np.array_equal(r.bound_offsets, boff)
    assert int(np.count_nonzero(r.status)) == 0 and int(np.count_nonzero(st)) == 0
    assert np.array_equal(r.scores, sc), "scores differ at %s" % np.nonzero(r.scores != sc)[0][:5]
    assert np.array_equal(r.boundaries, bd)
    assert np.array_equal(r.boundaries, (sc > 0).astype(np.uint8))
    if states:
        cs, ts, coff = o.predict_batch_states(text, offs, nthreads=NTHREADS)
        assert np.array_equal(r.char_offsets, coff)
        assert np.array_equal(r.char_states, cs), "char states differ at %s" % np.nonzero(r.char_states != cs)[0][:5]
        assert np.array_equal(r.type_states, ts), "type states differ at %s" % np.nonzero(r.type_states != ts)[0][:5]
    return r


def test_config2_full_size():
    """configs[1]: bccwj-suw-shaped, 300 000 char 1-3-gram patterns, no dictionary, no tags."""
    mb = synth.gen_model_bccwj_shaped(n_patterns=300_000, sample_sentences=200_000)
    p, o = vb.Predictor(vb.Model.read(mb)), OraclePredictor(mb)
    assert p.info["n_char_patterns"] == 300_000 and p.info["fast_path"] == 1
    text, offs, _ = _text(1)
    _compare_batch(p, o, text, offs)
    # ragged lengths (clipped log-normal): tiles of very different fill, groups split into ranges
    text, offs, _ = synth.gen_text(60_000, 40, seed=synth.TEXT_SEED + 2, ragged=True)
    _compare_batch(p, o, text, offs)


def test_config3_full_size_with_states():
    """configs[2]: + 20 000 tag models, predict_tags = true: scores, boundaries and the pattern-id states of both
    scorers for every character (what tag prediction consumes)."""
    mb = synth.gen_model_bccwj_shaped(n_patterns=300_000, sample_sentences=200_000, tag_models=20_000)
    p = vb.Predictor(vb.Model.read(mb), predict_tags=True)
    # the oracle's literal build-time merge of 20 000 tag models takes minutes; pattern ids and boundary weights do not
    # depend on it (oracle.py: states_only)
    o = OraclePredictor(mb, predict_tags=True, states_only=True)
    assert p.info["predict_tags"] == 1
    text, offs, _ = _text(3)
    _compare_batch(p, o, text, offs, states=True)


def test_config4_full_size():
    """configs[3]: KyTea-shaped, + 500 000 dictionary words (patterns longer than the window, rows outside the inline
    window: backward walk + overflow rows)."""
    mb = synth.gen_model_bccwj_shaped(n_patterns=300_000, sample_sentences=200_000, dict_words=500_000)
    p, o = vb.Predictor(vb.Model.read(mb)), OraclePredictor(mb)
    text, offs, _ = _text(4)
    _compare_batch(p, o, text, offs)


def test_fill_tags_synthetic_tag_model():
    """fill_tags through the GPU's states on a synthetic tag model (1 500 tag models): token strings with tags, sentence
    by sentence, against the oracle's predict_tags + write_tokenized_text."""
    mb = synth.gen_model_bccwj_shaped(n_patterns=30_000, sample_sentences=50_000, tag_models=1_500)
    p, o = vb.Predictor(vb.Model.read(mb), predict_tags=True), OraclePredictor(mb, predict_tags=True)
    text, offs, _ = synth.gen_text(1_500, 40, seed=synth.TEXT_SEED + 5)
    n_tagged = 0
    for i in range(len(offs) - 1):
        raw = bytes(text[int(offs[i]):int(offs[i + 1])]).decode()
        s = vb.Sentence.from_raw(raw)
        p.predict(s)
        s.fill_tags()
        got = s.write_tokenized_text()
        want = o.tokenize(raw, fill_tags=True)
        assert got == want, (i, raw)
        n_tagged += got.count("/")
    assert n_tagged > 100  # the synthetic tag models do fire


def test_config3_full_size_device_tags():
    """configs[2] with tag prediction on the device (k_tags): 20 000 tag models, the whole 120 000-sentence batch in one
    call; tag_token / tag_cand compared with the host restatement (vpt_fill_tags, itself pinned to the oracle and the
    reference's known answers at sizes the oracle can build) on a sample of sentences."""
    mb = synth.gen_model_bccwj_shaped(n_patterns=300_000, sample_sentences=200_000, tag_models=20_000)
    p = vb.Predictor(vb.Model.read(mb), predict_tags=True)
    text, offs, _ = _text(6)
    res, tok, cand, unserved = p.predict_batch_tags(text, offs)
    assert unserved == 0 and int(np.count_nonzero(res.status)) == 0
    assert int((tok >= 0).sum()) > 100_000      # the synthetic tokens are frequent strings: most tokens are known
    for i in range(0, N_SENT, N_SENT // 300):
        raw = bytes(text[int(offs[i]):int(offs[i + 1])]).decode()
        hs = vb.Sentence.from_raw(raw)
        p.predict(hs)
        hs.fill_tags()
        c0, c1 = int(res.char_offsets[i]), int(res.char_offsets[i + 1])
        assert tok[c0:c1].tolist() == hs._tag_token.tolist(), i
        assert cand[c0:c1].reshape(-1).tolist() == hs._tag_cand.reshape(-1).tolist(), i
