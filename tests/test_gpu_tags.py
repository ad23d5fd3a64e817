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
