"""Pins the CPU oracle (oracle/) to every known-answer test the reference holds for the
predict path.  CPU only."""
import os

import numpy as np
import pytest

from golden import reference_kat as kat
from vpt_testlib.bincode_model import encode_model
from vpt_testlib.oracle import OracleError, OraclePredictor, char_types

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def read(fn):
    with open(os.path.join(GOLDEN, fn), "rb") as f:
        return f.read()


@pytest.mark.parametrize("name", sorted(kat.SCORE_CASES))
def test_reference_score_vectors(name):
    case = kat.SCORE_CASES[name]
    p = OraclePredictor(encode_model(case["model"]), predict_tags=False)
    scores, bounds = p.predict(case["text"])
    assert scores.tolist() == case["scores"]
    if "boundaries" in case:
        assert bounds.tolist() == case["boundaries"]
    else:
        assert bounds.tolist() == [1 if s > 0 else 0 for s in case["scores"]]
    if "type_variant" in case:
        assert p.type_variant == case["type_variant"]


@pytest.mark.parametrize("name", sorted(kat.TAG_SCORE_CASES))
def test_reference_score_vectors_tag_variants(name):
    case = kat.TAG_SCORE_CASES[name]
    p = OraclePredictor(encode_model(case["model"]), predict_tags=True)
    scores, bounds = p.predict(case["text"])
    assert scores.tolist() == case["scores"]
    for (token_id, pos), want in case.get("tag_scores", []):
        which = 0 if "char" in name else 1
        got = p.add_tag_scores(which, case["text"], token_id, pos, [1] * 8)
        assert got.tolist() == want


def test_reference_predict_tags():
    # predictor.rs:863-903
    case = kat.PREDICT_BOUNDARIES
    p = OraclePredictor(encode_model(case["model"]), predict_tags=True)
    assert p.n_tags == 2
    tt, ti = p.predict_tags(case["text"])
    tags = []
    tm = case["model"]["tag_models"]
    for i in range(len(tt)):
        for k in range(2):
            if tt[i] < 0 or ti[i, k] < 0:
                tags.append(None)
            else:
                tags.append(tm[tt[i]]["tags"][k][ti[i, k]])
    assert tags == case["tags"]


def test_fill_tags_unsupported():
    # predictor.rs:974-1000: fill_tags on a predictor built with predict_tags=false panics
    p = OraclePredictor(encode_model(kat.PREDICTOR_TEST_MODEL), predict_tags=False)
    with pytest.raises(OracleError):
        p.predict_tags("この人は地球人だ")


def test_weight_merger_char():
    # char_scorer.rs:171-185.  offsets: n-gram => -window(3); "京都" second add has offset -2 = -len("京都")
    # => a dictionary word; "大阪" offset -2 likewise.
    m = dict(char_ngrams=[("東京都", [1, 2, 3, 4]), ("京都", [2, 4, 6, 8, 10])],
             dict=[("京都", [3, 6, 9], ""), ("大阪", [4, 8, 12], "")], bias=0, char_window=3, type_window=0)
    p = OraclePredictor(encode_model(m))
    got = [(pat.decode(), off, w) for pat, off, w in p.dump_patterns(0)]
    assert got == [(a, b, c) for a, b, c in kat.CHAR_WEIGHT_MERGER["merged"]]


def test_weight_merger_type():
    # type_scorer.rs:194-208 (all offsets -3; "cd" has -2 in the reference test, which no whole model can
    # produce for a type n-gram, so it is checked with offset -3 and the same weights).
    m = dict(type_ngrams=[(b"\x05\x01\x02", [1, 2, 3, 4]), (b"\x01\x02", [2, 4, 6, 8, 10]), (b"\x03\x04", [4, 8, 12])],
             bias=0, char_window=0, type_window=4)
    # duplicates are impossible through TypeScorerBoundary::new only if the list has them; add one:
    m["type_ngrams"].insert(2, (b"\x01\x02", [3, 6, 9]))
    # window 4 => offsets -4 (uniform shift does not change the merge arithmetic)
    p = OraclePredictor(encode_model(m))
    got = {pat: (off, w) for pat, off, w in p.dump_patterns(1)}
    assert got[b"\x01\x02"] == (-4, [5, 10, 15, 8, 10])
    assert got[b"\x03\x04"] == (-4, [4, 8, 12])
    assert got[b"\x05\x01\x02"] == (-4, [6, 12, 18, 12, 10])


def test_positional_weight_add_assign():
    # predictor.rs:678-747, all eight alignment cases
    import ctypes as C
    from vpt_testlib.oracle import lib
    L = lib()
    L.ora_pw_add.restype = C.c_long
    for (oy, y), (ox, x), (oz, z) in kat.POSITIONAL_WEIGHT_ADD:
        ya = np.zeros(32, np.int32)
        ya[: len(y)] = y
        xa = np.array(x, np.int32)
        off = C.c_int(oy)
        n = L.ora_pw_add(C.byref(off), C.c_void_p(ya.ctypes.data), C.c_size_t(len(y)), C.c_size_t(32), C.c_int(ox),
                         C.c_void_p(xa.ctypes.data), C.c_size_t(len(x)))
        assert (off.value, ya[:n].tolist()) == (oz, z)


def test_model_bin_fixture():
    data = read("model.bin")
    p = OraclePredictor(data, predict_tags=True)
    assert p.consumed == len(data) == 394
    for text, tags, want in kat.MODEL_BIN_TOKENIZE:
        assert p.tokenize(text, fill_tags=tags) == want
    p2 = OraclePredictor(data, predict_tags=False)
    for text, want in kat.MODEL_BIN_SCORES.items():
        assert p2.predict(text)[0].tolist() == want
        assert p.predict(text)[0].tolist() == want
    for text, tags, want in kat.MODEL_BIN_TOKENIZE:
        if not tags:
            assert p2.tokenize(text) == want


def test_docs_tok_fixture():
    # BASELINE config 1: resources/docs.tok with resources/model.bin
    p = OraclePredictor(read("model.bin"), predict_tags=True)
    for line in read("docs.tok").decode().splitlines():
        raw = "".join(tok.split("/")[0] for tok in line.split(" "))
        assert p.tokenize(raw, fill_tags=True) == line


def test_tantivy_fixture():
    data = read("tantivy_model.bin")
    p = OraclePredictor(data)
    assert p.consumed == len(data)
    for text, want in kat.TANTIVY_TOKENIZE:
        assert p.tokenize(text) == want


def test_char_types():
    # sentence.rs:50-67 and the doc example at sentence.rs:975-990
    assert char_types("A1あエ漢?").tolist() == [2, 1, 3, 4, 5, 6]
    assert char_types("Ａ１ｱ").tolist() == [2, 1, 4]
    assert char_types("぀ゖ゗゠ヺ・ーヿ").tolist() == [3, 3, 6, 4, 4, 6, 4, 4]
    assert char_types("\U00020000\U0002a6df\U0002a6e0\U0002f800\U0002fa1f\U0002fa20").tolist() == [5, 5, 6, 5, 5, 6]


def test_invalid_inputs():
    p = OraclePredictor(read("model.bin"))
    with pytest.raises(OracleError) as e:
        p.predict("")
    assert "at least one character" in str(e.value)
    with pytest.raises(OracleError) as e:
        p.predict("a\0b")
    assert "NULL" in str(e.value)
    s, b = p.predict("あ")
    assert len(s) == 0 and len(b) == 0
    with pytest.raises(OracleError):
        OraclePredictor(b"VaporettoTokenizer 0.4.0\n" + b"\0" * 16)


def test_batch_matches_single():
    p = OraclePredictor(read("model.bin"))
    sents = ["まぁ社長は火星猫だ", "まぁ良いだろう", "猫", "火星猫", ""]
    blob = "".join(sents).encode()
    offs = np.zeros(len(sents) + 1, np.uint64)
    np.cumsum([len(s.encode()) for s in sents], out=offs[1:])
    text = np.frombuffer(blob, np.uint8)
    for nt in (1, 3):
        scores, bounds, boff, status = p.predict_batch(text, offs, nthreads=nt)
        assert status.tolist() == [0, 0, 0, 0, 2]
        for i, s in enumerate(sents[:-1]):
            want = p.predict(s)[0]
            assert scores[int(boff[i]):int(boff[i + 1])].tolist() == want.tolist()


def test_char_types_every_code_point():
    """CharacterType::get_type (sentence.rs:50-67) over the whole code space: the oracle's restatement and the library's
    host function (vpt_char_types: the arithmetic form the kernels use) against the reference's ranges, listed as the
    reference lists them."""
    import numpy as np
    import vaporetto_b200 as vb
    ranges = [(1, [(0x30, 0x39), (0xFF10, 0xFF19)]),
              (2, [(0x41, 0x5A), (0x61, 0x7A), (0xFF21, 0xFF3A), (0xFF41, 0xFF5A)]),
              (3, [(0x3040, 0x3096)]),
              (4, [(0x30A0, 0x30FA), (0x30FC, 0x30FF), (0xFF66, 0xFF9F)]),
              (5, [(0x3400, 0x4DBF), (0x4E00, 0x9FFF), (0xF900, 0xFAFF), (0x20000, 0x2A6DF), (0x2A700, 0x2B73F),
                   (0x2B740, 0x2B81F), (0x2B820, 0x2CEAF), (0x2F800, 0x2FA1F)])]
    want = np.full(0x110000, 6, np.uint8)
    for ty, rs in ranges:
        for lo, hi in rs:
            want[lo:hi + 1] = ty
    cps = [c for c in range(1, 0x110000) if not 0xD800 <= c <= 0xDFFF]
    for lo in range(0, len(cps), 65536):
        chunk = cps[lo:lo + 65536]
        text = "".join(map(chr, chunk))
        assert char_types(text).tolist() == want[chunk].tolist()
        assert vb.Sentence.from_raw(text).char_types().tolist() == want[chunk].tolist()
