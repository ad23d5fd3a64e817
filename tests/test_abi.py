"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/vaporetto_b200.h declares, parses models on the host, and refuses to compute without a GPU."""
import os
import re

import numpy as np
import pytest

import vaporetto_b200 as vb
from golden import reference_kat as kat
from vpt_testlib.bincode_model import encode_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def header_symbols():
    src = open(os.path.join(ROOT, "include", "vaporetto_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vpt_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = vb.lib()
    names = header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/vaporetto_b200.h but not exported"
    assert sorted(n for n, _, _ in vb.ABI) == names
    assert b"sm_100a" in L.vpt_version()


def test_model_read_on_host():
    data = open(os.path.join(GOLDEN, "model.bin"), "rb").read()
    m, rest = vb.Model.read_slice(data + b"xyz")
    assert m.consumed == 394 and rest == b"xyz"
    m2 = vb.Model.read(encode_model(kat.PREDICTOR_TEST_MODEL))
    assert m2.consumed > 0


def test_model_read_errors():
    with pytest.raises(vb.VaporettoError) as e:
        vb.Model.read(b"VaporettoTokenizer 0.4.0\n" + b"\0" * 8)
    assert e.value.kind == "InvalidModel" and "model version mismatch" in str(e.value)
    good = open(os.path.join(GOLDEN, "model.bin"), "rb").read()
    with pytest.raises(vb.VaporettoError) as e:
        vb.Model.read(good[:200])
    assert e.value.kind == "DecodeError"


def test_sentence_host_helpers():
    s = vb.Sentence.from_raw("A1あエ漢?")
    assert s.char_types().tolist() == [2, 1, 3, 4, 5, 6]
    assert s.boundaries().tolist() == [2] * 5
    with pytest.raises(vb.VaporettoError) as e:
        vb.Sentence.from_raw("")
    assert "must contain at least one character" in str(e.value)
    with pytest.raises(vb.VaporettoError) as e:
        vb.Sentence.from_raw("a\0b")
    assert "must not contain NULL" in str(e.value)
    s = vb.Sentence.from_raw("まぁ社長は火星猫だ")
    with pytest.raises(vb.VaporettoError):
        s.update_raw("")
    assert s.as_raw_text() == " "  # sentence.rs:264-283: replaced with a white space on error
    # write_tokenized_text with hand-set boundaries (sentence.rs:850-886 doc example, escaping)
    s = vb.Sentence.from_raw("a/b c")
    s.boundaries_mut()[:] = [0, 0, 1, 0]
    assert s.write_tokenized_text() == "a\\/b \\ c"
    s.boundaries_mut()[:] = [0, 2, 1, 0]
    assert s.write_tokenized_text() == "\\ c"
    assert [t.surface() for t in s.iter_tokens()] == [" c"]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(vb.VaporettoError) as e:
        vb.Predictor(vb.Model.read(open(os.path.join(GOLDEN, "model.bin"), "rb").read()))
    assert e.value.kind == "CudaError"


# ---- host-side tag prediction (vpt_fill_tags) on CPU, fed with the oracle's boundaries and states -------------

def _host_tags(model_bytes, text):
    from vpt_testlib.oracle import OraclePredictor
    o = OraclePredictor(model_bytes, predict_tags=True)
    _, bd, cs, ts = o.predict(text, states=True)
    p = vb.Predictor(vb.Model.read(model_bytes), predict_tags=True, device=-1)  # host-only handle
    s = vb.Sentence.from_raw(text)
    s._boundaries = bd.astype(np.uint8)
    s._char_states, s._type_states, s._predictor = cs, ts, p
    s.fill_tags()
    tt, ti = o.predict_tags(text)
    return s, p, o, tt, ti


def test_fill_tags_reference_vector():
    # predictor.rs:863-903 test_predict_tags
    case = kat.PREDICT_BOUNDARIES
    s, p, o, tt, ti = _host_tags(encode_model(case["model"]), case["text"])
    assert s.n_tags() == 2 and s.tags() == case["tags"]
    assert s._tag_token.tolist() == tt.tolist()
    assert s._tag_cand.reshape(-1, 2).tolist() == ti.tolist()


def test_fill_tags_fixture_models():
    data = open(os.path.join(GOLDEN, "model.bin"), "rb").read()
    for text, tags, want in kat.MODEL_BIN_TOKENIZE:
        if not tags:
            continue
        s, p, o, tt, ti = _host_tags(data, text)
        assert s.write_tokenized_text() == want


def test_host_only_predictor_cannot_score():
    p = vb.Predictor(vb.Model.read(open(os.path.join(GOLDEN, "model.bin"), "rb").read()), device=-1)
    with pytest.raises(vb.VaporettoError) as e:
        p.predict(vb.Sentence.from_raw("猫"))
    assert e.value.kind == "CudaError" and "no CPU fallback" in str(e.value)


def test_fill_tags_random_models():
    from hypothesis import given, settings, strategies as st
    alpha = "あいう人火星aB1"
    chars = st.sampled_from(list(alpha))
    w = st.integers(-50, 50)

    @st.composite
    def tag_model(draw, i, cw, tw):
        ncand = draw(st.lists(st.integers(1, 4), min_size=1, max_size=3))
        slen = sum(c for c in ncand if c >= 2)
        cn = [("".join(draw(st.lists(chars, min_size=1, max_size=3))),
               [(draw(st.integers(0, cw)), draw(st.lists(w, min_size=1, max_size=slen + 2)))]) for _ in range(draw(st.integers(0, 5)))]
        tn = [(bytes(draw(st.lists(st.integers(1, 6), min_size=1, max_size=3))),
               [(draw(st.integers(0, tw)), draw(st.lists(w, min_size=1, max_size=slen + 2)))]) for _ in range(draw(st.integers(0, 3)))]
        return dict(token="".join(draw(st.lists(chars, min_size=1, max_size=2))) + str(i) * 0,
                    tags=[["t%d_%d" % (k, j) for j in range(c)] for k, c in enumerate(ncand)],
                    char_ngrams=cn, type_ngrams=tn, bias=draw(st.lists(w, min_size=slen, max_size=slen + 3)))

    @st.composite
    def models(draw):
        cw, tw = draw(st.integers(1, 3)), draw(st.integers(1, 4))
        cng = {"".join(draw(st.lists(chars, min_size=1, max_size=3))): draw(st.lists(w, min_size=1, max_size=4)) for _ in range(4)}
        tng = {bytes(draw(st.lists(st.integers(1, 6), min_size=1, max_size=2))): draw(st.lists(w, min_size=1, max_size=4)) for _ in range(3)}
        tms = [draw(tag_model(i, cw, tw)) for i in range(draw(st.integers(1, 4)))]
        seen, uniq = set(), []
        for t in tms:
            if t["token"] not in seen:
                seen.add(t["token"])
                uniq.append(t)
        return dict(char_ngrams=list(cng.items()), type_ngrams=list(tng.items()), dict=[("人", [30, -30], "")], bias=3,
                    char_window=cw, type_window=tw, tag_models=uniq)

    @settings(max_examples=150, deadline=None)
    @given(models(), st.lists(chars, min_size=1, max_size=16))
    def run(model, text):
        text = "".join(text)
        s, p, o, tt, ti = _host_tags(encode_model(model), text)
        assert s._tag_token.tolist() == tt.tolist()
        assert s._tag_cand.reshape(len(tt), -1).tolist() == ti.tolist()

    run()
