"""Cross-checks the C++ oracle against the independent brute-force scorer on the reference's
known-answer vectors and on hypothesis-generated models."""
import os
import sys

import pytest
from hypothesis import given, settings, strategies as st

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import bruteforce  # noqa: E402

from golden import reference_kat as kat  # noqa: E402
from vpt_testlib.bincode_model import encode_model  # noqa: E402
from vpt_testlib.oracle import OraclePredictor  # noqa: E402


@pytest.mark.parametrize("name", sorted(kat.SCORE_CASES))
def test_bruteforce_on_reference_vectors(name):
    case = kat.SCORE_CASES[name]
    scores, _ = bruteforce.predict(case["model"], case["text"])
    assert scores == case["scores"]


ALPHA = "あいうアイ人火星地球aB1。"
chars = st.sampled_from(list(ALPHA))
w = st.integers(-40000, 40000)


@st.composite
def models(draw):
    cw = draw(st.integers(0, 5))
    tw = draw(st.integers(0, 5))
    cng = {}
    for _ in range(draw(st.integers(0, 8))):
        g = "".join(draw(st.lists(chars, min_size=1, max_size=4)))
        L = len(g)
        full = max(2 * cw - L + 1, 0)
        cng[g] = draw(st.lists(w, min_size=0, max_size=full + 1))
    dic = []
    for _ in range(draw(st.integers(0, 6))):
        g = "".join(draw(st.lists(chars, min_size=1, max_size=9)))
        dic.append((g, draw(st.lists(w, min_size=0, max_size=len(g) + 2)), ""))
    tng = {}
    for _ in range(draw(st.integers(0, 6))):
        g = bytes(draw(st.lists(st.integers(1, 6), min_size=1, max_size=4)))
        full = max(2 * tw - len(g) + 1, 0)
        tng[g] = draw(st.lists(w, min_size=0, max_size=full + 1))
    bias = draw(st.sampled_from([0, 5, -7, 2**31 - 1, -2**31]))
    return dict(char_ngrams=list(cng.items()), type_ngrams=list(tng.items()), dict=dic, bias=bias,
                char_window=cw, type_window=tw)


@settings(max_examples=300, deadline=None)
@given(models(), st.lists(chars, min_size=1, max_size=24))
def test_oracle_equals_bruteforce(model, text):
    text = "".join(text)
    p = OraclePredictor(encode_model(model))
    got, gb = p.predict(text)
    want, wb = bruteforce.predict(model, text)
    assert got.tolist() == want
    assert gb.tolist() == wb


def test_oracle_double_array_walk_equals_hash_walk(monkeypatch):
    """The oracle walks the automaton as a double array (the layout of the reference's daachorse matcher: what the
    timed CPU arm runs); ORA_AC=hash keeps the hash-probed goto function it is built from.  Same transitions: scores,
    boundaries and pattern-id states agree on a synthetic model with n-grams, dictionary words and tags."""
    import numpy as np
    from vpt_testlib import synth

    mb = synth.gen_model_bccwj_shaped(n_patterns=20000, sample_sentences=50000, dict_words=30000)
    text, offs, _ = synth.gen_text(3000, 40)
    outs = []
    for mode in ("hash", "da"):
        monkeypatch.setenv("ORA_AC", mode)
        o = OraclePredictor(mb)
        sc, bd, boff, st_ = o.predict_batch(text, offs, nthreads=2)
        outs.append((sc.copy(), bd.copy(), boff.copy(), st_.copy()))
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
    mt = synth.gen_model_bccwj_shaped(n_patterns=5000, sample_sentences=20000, tag_models=300)
    souts = []
    for mode in ("hash", "da"):
        monkeypatch.setenv("ORA_AC", mode)
        o = OraclePredictor(mt, predict_tags=True)
        souts.append([np.copy(x) for x in o.predict_batch_states(text[:int(offs[500])], offs[:501], nthreads=1)])
    for a, b in zip(*souts):
        assert np.array_equal(a, b)
