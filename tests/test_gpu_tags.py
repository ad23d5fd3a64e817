"""Device-side tag prediction (k_tags, vpt_predict_batch_tags) against the host restatement (vpt_fill_tags) and the
CPU oracle's predict_tags (reference predictor.rs:546-637, known answers :863-903, tag scorers char_scorer.rs:405-525 /
type_scorer.rs:367-473).  tag_token / tag_cand must be identical for every character."""
import os

import numpy as np
import pytest

import vaporetto_b200 as vb
from vpt_testlib import synth
from vpt_testlib.bincode_model import encode_model
from vpt_testlib.oracle import OraclePredictor
from golden import reference_kat as kat
from test_gpu_parity import _random_model, read

pytestmark = pytest.mark.gpu


def _batch(sents):
    enc = [s.encode() for s in sents]
    offs = np.zeros(len(enc) + 1, np.uint64)
    np.cumsum([len(e) for e in enc], out=offs[1:])
    return np.frombuffer(b"".join(enc), np.uint8), offs


def _check_against_host_and_oracle(p, o, sents):
    text, offs = _batch(sents)
    res, tok, cand, unserved = p.predict_batch_tags(text, offs)
    assert unserved == 0
    nt = p.n_tags
    for i, s in enumerate(sents):
        c0, c1 = int(res.char_offsets[i]), int(res.char_offsets[i + 1])
        # host path: predict + fill_tags of the Python mirror (vpt_fill_tags)
        hs = vb.Sentence.from_raw(s)
        p.predict(hs)
        hs.fill_tags()
        assert tok[c0:c1].tolist() == hs._tag_token.tolist(), (i, s)
        assert cand[c0:c1].reshape(-1).tolist() == hs._tag_cand.reshape(-1).tolist(), (i, s)
        # oracle: token ids may be numbered differently only if tokens repeat in the model; candidates must agree
        ott, oti = o.predict_tags(s)
        assert (tok[c0:c1] >= 0).tolist() == (ott >= 0).tolist(), (i, s)
        assert cand[c0:c1].tolist() == oti.tolist(), (i, s)
    return res


def test_reference_known_answers():
    # predictor.rs:863-903 (tags of "この人は地球人だ") and the bundled model's doctest sentences
    mb = encode_model(kat.PREDICTOR_TEST_MODEL)
    p, o = vb.Predictor(vb.Model.read(mb), predict_tags=True), OraclePredictor(mb, predict_tags=True)
    _check_against_host_and_oracle(p, o, ["この人は地球人だ", "地球人", "この人"])
    mb = read("model.bin")
    p, o = vb.Predictor(vb.Model.read(mb), predict_tags=True), OraclePredictor(mb, predict_tags=True)
    _check_against_host_and_oracle(p, o, ["まぁ社長は火星猫だ", "まぁ良いだろう", "火星", "社長は社長だ" * 30])


@pytest.mark.parametrize("cw,tw,maxdict,tags", [(3, 3, 5, 3), (1, 5, 3, 2), (2, 2, 2, 4), (3, 3, 9, 6)])
def test_random_tag_models(cw, tw, maxdict, tags):
    rng = np.random.default_rng(77 + 1000 * cw + 100 * tw + maxdict + tags)
    for _ in range(3):
        model, alpha = _random_model(rng, cw, tw, maxdict=maxdict, tags=tags)
        mb = encode_model(model)
        p, o = vb.Predictor(vb.Model.read(mb), predict_tags=True), OraclePredictor(mb, predict_tags=True)
        sents = ["".join(rng.choice(list(alpha), size=rng.integers(1, 60))) for _ in range(300)]
        sents += ["".join(rng.choice(list(alpha), size=n)) for n in (1, 2, 31, 32, 33, 64, 65, 300)]
        _check_against_host_and_oracle(p, o, sents)


def test_synthetic_tag_model_batch():
    """1 500 tag models on a 30 000-pattern model, 3 000 sentences: the arrays of the whole batch against fill_tags."""
    mb = synth.gen_model_bccwj_shaped(n_patterns=30_000, sample_sentences=50_000, tag_models=1_500)
    p, o = vb.Predictor(vb.Model.read(mb), predict_tags=True), OraclePredictor(mb, predict_tags=True)
    text, offs, _ = synth.gen_text(3_000, 40, seed=synth.TEXT_SEED + 11)
    sents = [bytes(text[int(offs[i]):int(offs[i + 1])]).decode() for i in range(len(offs) - 1)]
    res = _check_against_host_and_oracle(p, o, sents[:400])
    # the whole batch in one call: every known token gets a token id, and scores / boundaries are the plain batch's
    r2, tok, cand, unserved = p.predict_batch_tags(text, offs)
    assert unserved == 0 and int((tok >= 0).sum()) > 1000
    plain = p.predict_batch(text, offs)
    assert np.array_equal(r2.scores, plain.scores) and np.array_equal(r2.boundaries, plain.boundaries)


def test_predict_tags_false_is_rejected():
    mb = read("model.bin")
    p = vb.Predictor(vb.Model.read(mb), predict_tags=False)
    text, offs = _batch(["まぁ社長は火星猫だ"])
    with pytest.raises(vb.VaporettoError):
        p.predict_batch_tags(text, offs)


# ---- vpt_predict_batch_compact: one bit per boundary, one record per token -------------------------------------------------

def _check_compact(p, text, offs, tags, monkeypatch=None):
    plain = p.predict_batch(text, offs)
    r = p.predict_batch_compact(text, offs, tags=tags)
    n = len(offs) - 1
    assert r.n_boundaries == plain.boundaries.size
    assert np.array_equal(r.boundaries(), plain.boundaries)
    assert np.array_equal(r.status.astype(np.int32), plain.status)
    nb = np.diff(plain.bound_offsets.astype(np.int64))
    # a sentence owns n_chars - 1 bits whether it was scored or rejected (a rejected sentence's bits are 0)
    assert np.array_equal(np.maximum(r.n_chars.astype(np.int64) - 1, 0), nb)
    ok = plain.status == 0
    assert np.array_equal(r.n_chars.astype(np.int64)[ok], nb[ok] + 1)
    # tokens per sentence = boundaries set + 1 (0 for a rejected sentence)
    csum = np.concatenate(([0], np.cumsum(plain.boundaries.astype(np.int64))))
    ones = csum[plain.bound_offsets[1:].astype(np.int64)] - csum[plain.bound_offsets[:-1].astype(np.int64)]
    assert np.array_equal(r.n_tokens.astype(np.int64), np.where(plain.status == 0, ones + 1, 0))
    for s in (0, n // 2, n - 1):
        lo, hi = int(plain.bound_offsets[s]), int(plain.bound_offsets[s + 1])
        assert np.array_equal(r.boundaries(s), plain.boundaries[lo:hi])
    if not tags:
        assert r.token_ids is None
        return r
    res, tok, cand, unserved = p.predict_batch_tags(text, offs)
    assert unserved == 0 and r.n_unserved == 0
    # the per-character arrays hold a token's result at its last character: compacting them gives the records
    ends = np.ones(tok.size, bool)
    for s in range(n):
        c0, c1 = int(res.char_offsets[s]), int(res.char_offsets[s + 1])
        b0 = int(res.bound_offsets[s])
        if c1 > c0:
            ends[c0:c1 - 1] = res.boundaries[b0:b0 + (c1 - c0 - 1)] == 1
        if res.status[s] != 0:
            ends[c0:c1] = False  # a rejected sentence has no tokens
    assert r.token_ids.size == int(ends.sum()) == int(r.n_tokens.sum())
    assert np.array_equal(r.token_ids, tok[ends])
    want_c = cand[ends]
    assert np.array_equal(r.token_cands.astype(np.int32), np.where(want_c < 0, 255, want_c))
    return r


def test_compact_reference_known_answers():
    mb = encode_model(kat.PREDICTOR_TEST_MODEL)
    p = vb.Predictor(vb.Model.read(mb), predict_tags=True)
    text, offs = _batch(["この人は地球人だ", "地球人", "この人", "人"])
    r = _check_compact(p, text, offs, tags=True)
    # predictor.rs:863-903: tokens of "この人は地球人だ" with their tags through the token records
    toks = []
    s = vb.Sentence.from_raw("この人は地球人だ")
    p.predict(s)
    s.fill_tags()
    want = [[t.tags()[k] for k in range(p.n_tags)] for t in s.iter_tokens()]
    lo, hi = int(r.token_offsets[0]), int(r.token_offsets[1])
    for rec in range(lo, hi):
        tid = int(r.token_ids[rec])
        toks.append([None if tid < 0 or r.token_cands[rec, k] == 255 else p.tag_string(tid, k, int(r.token_cands[rec, k]))
                     for k in range(p.n_tags)])
    assert toks == want


@pytest.mark.parametrize("chunk", [None, "1024", "5000"])
def test_compact_synthetic_batch(chunk, monkeypatch):
    """Bit stream, per-sentence words and token records of 20 000 sentences against the per-character interfaces, with
    chunk sizes that put chunk borders inside bit words (VPT_CHUNK_SENTENCES is read once per process: the small
    sizes run in a child process)."""
    if chunk is not None:
        import subprocess
        import sys
        code = ("import os,sys; sys.path.insert(0, %r); os.environ['VPT_CHUNK_SENTENCES']=%r\n"
                "import test_gpu_tags as t; t._compact_synthetic()\n") % (os.path.dirname(os.path.abspath(__file__)), chunk)
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr[-2000:]
        return
    _compact_synthetic()


def _compact_synthetic():
    mb = synth.gen_model_bccwj_shaped(n_patterns=30_000, sample_sentences=50_000, tag_models=1_500)
    p = vb.Predictor(vb.Model.read(mb), predict_tags=True)
    text, offs, _ = synth.gen_text(20_000, 40, seed=synth.TEXT_SEED + 12)
    # a few sentences the device rejects (empty, NUL) and single-character sentences
    sents = [bytes(text[int(offs[i]):int(offs[i + 1])]) for i in range(len(offs) - 1)]
    sents[5] = b""
    sents[77] = b"a\x00b"
    sents[78] = "あ".encode()
    sents[4999] = b""
    enc = b"".join(sents)
    offs2 = np.zeros(len(sents) + 1, np.uint64)
    np.cumsum([len(e) for e in sents], out=offs2[1:])
    t2 = np.frombuffer(enc, np.uint8)
    r = _check_compact(p, t2, offs2, tags=True)
    assert int((r.token_ids >= 0).sum()) > 1000
    _check_compact(p, t2, offs2, tags=False)
    p0 = vb.Predictor(vb.Model.read(mb), predict_tags=False)
    _check_compact(p0, t2, offs2, tags=False)
    with pytest.raises(vb.VaporettoError):
        p0.predict_batch_compact(t2, offs2, tags=True)
