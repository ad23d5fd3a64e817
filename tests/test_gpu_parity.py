"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle and the reference's
known-answer vectors.  Bit-exact (integer work).  Run with `-m gpu` on a B200."""
import os

import numpy as np
import pytest

import vaporetto_b200 as vb
from golden import reference_kat as kat
from vpt_testlib import synth
from vpt_testlib.bincode_model import encode_model
from vpt_testlib.oracle import OraclePredictor

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def read(fn):
    with open(os.path.join(GOLDEN, fn), "rb") as f:
        return f.read()


def make(model_bytes, tags=False):
    return vb.Predictor(vb.Model.read(model_bytes), predict_tags=tags)


# ---- the reference's own unit tests, through the mirrored API ------------------------------------------

@pytest.mark.parametrize("name", sorted(kat.SCORE_CASES))
def test_reference_score_vectors(name):
    case = kat.SCORE_CASES[name]
    p = make(encode_model(case["model"]))
    s = vb.Sentence.from_raw(case["text"])
    p.predict(s)
    assert s.boundary_scores().tolist() == case["scores"]
    want_b = case.get("boundaries", [1 if x > 0 else 0 for x in case["scores"]])
    assert s.boundaries().tolist() == want_b


@pytest.mark.parametrize("name", sorted(kat.TAG_SCORE_CASES))
def test_reference_score_vectors_tag_variants(name):
    case = kat.TAG_SCORE_CASES[name]
    mb = encode_model(case["model"])
    p = make(mb, tags=True)
    s = vb.Sentence.from_raw(case["text"])
    p.predict(s)
    assert s.boundary_scores().tolist() == case["scores"]
    o = OraclePredictor(mb, predict_tags=True)
    _, _, ocs, ots = o.predict(case["text"], states=True)
    if p.info["char_scorer"] == 2:
        assert s._char_states.tolist() == ocs.tolist()
    if p.info["type_scorer"] == 3:
        assert s._type_states.tolist() == ots.tolist()


def test_reference_predict_tags():
    # predictor.rs:863-903 test_predict_tags
    case = kat.PREDICT_BOUNDARIES
    p = make(encode_model(case["model"]), tags=True)
    s = vb.Sentence.from_raw(case["text"])
    p.predict(s)
    s.fill_tags()
    assert s.boundary_scores().tolist() == case["scores"]
    assert s.boundaries().tolist() == case["boundaries"]
    assert s.n_tags() == 2
    assert s.tags() == case["tags"]


def test_fill_tags_unsupported():
    # predictor.rs:974-987: panics in the reference; an InvalidArgument error here
    p = make(encode_model(kat.PREDICTOR_TEST_MODEL), tags=False)
    s = vb.Sentence.from_raw("この人は地球人だ")
    p.predict(s)
    with pytest.raises(vb.VaporettoError) as e:
        s.fill_tags()
    assert "predict_tags = false" in str(e.value)


def test_model_bin_doctests():
    # lib.rs:17-41, predictor.rs:388-429, sentence.rs:1121-1137
    data = read("model.bin")
    for tags in (False, True):
        p = make(data, tags=tags)
        for text, with_tags, want in kat.MODEL_BIN_TOKENIZE:
            if with_tags and not tags:
                continue
            s = vb.Sentence.from_raw(text)
            p.predict(s)
            assert s.boundary_scores().tolist() == kat.MODEL_BIN_SCORES[text]
            if with_tags:
                s.fill_tags()
            assert s.write_tokenized_text() == want


def test_docs_tok_config1():
    # BASELINE config 1: resources/docs.tok with resources/model.bin
    p = make(read("model.bin"), tags=True)
    for line in read("docs.tok").decode().splitlines():
        raw = "".join(tok.split("/")[0] for tok in line.split(" "))
        s = vb.Sentence.from_raw(raw)
        p.predict(s)
        s.fill_tags()
        assert s.write_tokenized_text() == line


def test_tantivy_fixture():
    p = make(read("tantivy_model.bin"))
    for text, want in kat.TANTIVY_TOKENIZE:
        s = vb.Sentence.from_raw(text)
        p.predict(s)
        assert s.write_tokenized_text() == want
        # byte offsets of tokens (vaporetto_tantivy/src/lib.rs:183-191)
    s = vb.Sentence.from_raw("東京特許許可局")
    p.predict(s)
    toks = list(s.iter_tokens())
    assert [(t.start(), t.end()) for t in toks] == [(0, 2), (2, 4), (4, 6), (6, 7)]


# ---- batches against the oracle ----------------------------------------------------------------------

def oracle_batch(o, text, offs):
    return o.predict_batch(text, offs, nthreads=4)


def check_batch(p, o, text, offs, want_states=False):
    r = p.predict_batch(text, offs, want_states=want_states)
    sc, bd, boff, st = oracle_batch(o, text, offs)
    assert r.bound_offsets.tolist() == boff.tolist()
    ok = st == 0
    assert (r.status == 0).tolist() == ok.tolist()
    assert np.array_equal(r.scores, sc)
    assert np.array_equal(r.boundaries, bd)
    return r


def test_edge_cases():
    data = read("model.bin")
    p, o = make(data), OraclePredictor(data)
    sents = ["まぁ社長は火星猫だ", "", "猫", "a\0b", "まぁ良いだろう", "𠀋𠀋火星🤌🏿", "x" * 700, "火" * 1000 + "星人" * 333, " "]
    blob = b"".join(s.encode() for s in sents) + b"\xe3\x81" + b"ok" + b"\xff"
    lens = [len(s.encode()) for s in sents] + [2, 2, 1]
    offs = np.zeros(len(lens) + 1, np.uint64)
    np.cumsum(lens, out=offs[1:])
    text = np.frombuffer(blob, np.uint8)
    r = p.predict_batch(text, offs)
    assert r.status.tolist() == [0, 1, 0, 2, 0, 0, 0, 0, 0, 3, 0, 3]
    for i, s in enumerate(sents):
        if r.status[i] != 0:
            continue
        want, wb = o.predict(s)
        assert r.sentence_scores(i).tolist() == want.tolist(), i
        assert r.sentence_boundaries(i).tolist() == wb.tolist(), i
    # rejected sentences keep zero-filled slots: "a\0b" has 3 chars -> 2 slots
    assert r.sentence_scores(3).tolist() == [0, 0]
    assert int(r.bound_offsets[2] - r.bound_offsets[1]) == 0
    # single-sentence entry point reports the reference's errors
    for bad, msg in (("", "at least one character"), ("a\0b", "NULL")):
        with pytest.raises(vb.VaporettoError) as e:
            vb.Sentence.from_raw(bad)
        assert msg in str(e.value)


@pytest.mark.parametrize("kind", ["fast", "general_dict", "general_tags"])
def test_long_sentences_and_tile_splitting(kind):
    """Sentences longer than one tile (slots / bytes), groups that must be split into several ranges, and
    mixtures of tiny and huge sentences inside one 64-sentence group."""
    rng = np.random.default_rng(11)
    if kind == "fast":
        mb = synth.gen_model_bccwj_shaped(n_patterns=8000, sample_sentences=20000)
        tags = False
    elif kind == "general_dict":
        mb = synth.gen_model_bccwj_shaped(n_patterns=8000, sample_sentences=20000, dict_words=5000)
        tags = False
    else:
        mb = synth.gen_model_bccwj_shaped(n_patterns=8000, sample_sentences=20000, tag_models=200)
        tags = True
    p, o = make(mb, tags=tags), OraclePredictor(mb, predict_tags=tags)
    lens = [1, 2, 5000, 3, 40, 3100, 3064, 3065, 12, 20000, 1, 700, 2900, 2900, 64, 1] + list(rng.integers(1, 400, size=150))
    lens += [4200] * 3 + list(rng.integers(1, 80, size=70))
    cps = synth.gen_codepoints(len(lens), np.array(lens), seed=99)
    text, offs = synth.encode_utf8(cps, np.array(lens))
    r = check_batch(p, o, text, offs, want_states=tags)
    if tags:
        for i in (2, 9, 11):
            s = bytes(text[int(offs[i]):int(offs[i + 1])]).decode()
            _, _, ocs, ots = o.predict(s, states=True)
            c0 = int(r.char_offsets[i])
            assert np.array_equal(r.char_states[c0:c0 + len(ocs)], ocs)
            assert np.array_equal(r.type_states[c0:c0 + len(ots)], ots)
    # all-ASCII sentence of 13 000 bytes: bytes exceed the tile text buffer while slots do too
    big = ("abc 123 " * 1700)
    rb = p.predict_batch(np.frombuffer(big.encode(), np.uint8), np.array([0, len(big)], np.uint64))
    assert rb.scores.tolist() == o.predict(big)[0].tolist()


def test_utf8_validation_matches_strict_decoder():
    """status 3 <=> the bytes are not valid UTF-8 (Python's strict decoder = Rust's str validity rules:
    no overlongs, no surrogates, <= U+10FFFF); 2 <=> valid but contains NUL; 1 <=> empty."""
    p = make(read("model.bin"))
    rng = np.random.default_rng(7)
    pool = "aZ0 \x00é߿ࠀ਀퟿\ue000￿𐀀𠀋\U0010ffffあ漢カ".encode("utf-8", "surrogatepass")
    special = [b"\xc0\x80", b"\xc1\xbf", b"\xe0\x80\x80", b"\xe0\x9f\xbf", b"\xe0\xa0\x80", b"\xed\x9f\xbf",
               b"\xed\xa0\x80", b"\xed\xbf\xbf", b"\xf0\x80\x80\x80", b"\xf0\x8f\xbf\xbf", b"\xf0\x90\x80\x80",
               b"\xf4\x8f\xbf\xbf", b"\xf4\x90\x80\x80", b"\xf5\x80\x80\x80", b"\xff", b"\xfe", b"\x80", b"\xbf",
               b"\xe3\x81", b"\xe3", b"\xf0\x9f\xa4", b"\xc3", b"\xe3\x81\x82\x82", b"\xe3\x41\x81\x81", b"\xc3\xc3\xa9"]
    sents = []
    for _ in range(3000):
        kind = rng.integers(0, 4)
        if kind == 0:  # valid text
            s = "".join(rng.choice(list("aZ0 é߿ࠀ퟿\ue000￿𐀀𠀋\U0010ffffあ漢カ"), size=rng.integers(0, 40))).encode()
        elif kind == 1:  # valid with a mutation
            b = bytearray("".join(rng.choice(list("aé߿ࠀ𐀀あ漢カ"), size=rng.integers(1, 30))).encode())
            for _ in range(rng.integers(1, 3)):
                b[rng.integers(0, len(b))] = int(rng.integers(0, 256))
            s = bytes(b)
        elif kind == 2:  # random bytes biased to the interesting ranges
            s = bytes(rng.choice(list(pool) + [0x80, 0xBF, 0xC0, 0xC2, 0xE0, 0xED, 0xF0, 0xF4, 0xF5, 0xFF],
                                 size=rng.integers(0, 24)).astype(np.uint8))
        else:  # special sequences embedded in valid text, truncated at random
            s = "あ漢".encode() + special[rng.integers(0, len(special))] + "カa".encode()
            s = s[rng.integers(0, 4):len(s) - rng.integers(0, 4)]
        sents.append(s)
    sents += special + [b"", b"\x00", b"a\x00", "漢".encode() * 200 + b"\xe3\x81", b"\x81" + "あ".encode() * 50]
    lens = [len(s) for s in sents]
    offs = np.zeros(len(lens) + 1, np.uint64)
    np.cumsum(lens, out=offs[1:])
    text = np.frombuffer(b"".join(sents) + b"\0" * 8, np.uint8)
    r = p.predict_batch(text[: int(offs[-1])] if False else text, offs)
    for i, s in enumerate(sents):
        try:
            u = s.decode("utf-8")
            want = 1 if len(u) == 0 else (2 if "\0" in u else 0)
        except UnicodeDecodeError:
            want = 3
        assert int(r.status[i]) == want, (i, s, int(r.status[i]), want)


def _random_model(rng, cw, tw, n_ng=40, n_dict=20, maxdict=9, tags=0):
    alpha = "あいうえおアイウ人火星地球猫社長aB1。、"
    def word(lo, hi):
        return "".join(rng.choice(list(alpha), size=rng.integers(lo, hi + 1)))
    cng = {}
    for _ in range(n_ng):
        g = word(1, 3)
        cng[g] = rng.integers(-32767, 32768, size=max(2 * cw - len(g) + 1, 0)).tolist()
    dic = [(word(1, maxdict),) for _ in range(n_dict)]
    dic = [(w, rng.integers(-32767, 32768, size=len(w) + 1).tolist(), "") for (w,) in dic]
    tng = {}
    for _ in range(30):
        g = bytes(rng.integers(1, 7, size=rng.integers(1, 4)).tolist())
        tng[g] = rng.integers(-32767, 32768, size=max(2 * tw - len(g) + 1, 0)).tolist()
    tms = []
    for t in range(tags):
        cn = [(word(1, 3), [(int(rng.integers(0, cw + 1)), rng.integers(-99, 99, size=2).tolist())]) for _ in range(5)]
        tn = [(bytes(rng.integers(1, 7, size=rng.integers(1, 4)).tolist()),
               [(int(rng.integers(0, tw + 1)), rng.integers(-99, 99, size=2).tolist())]) for _ in range(3)]
        tms.append(dict(token=word(1, 2), tags=[["x", "y"]], char_ngrams=cn, type_ngrams=tn, bias=[1, 2]))
    return dict(char_ngrams=list(cng.items()), type_ngrams=list(tng.items()), dict=dic,
                bias=int(rng.integers(-1000, 1000)), char_window=cw, type_window=tw, tag_models=tms), alpha


@pytest.mark.parametrize("cw,tw,maxdict,tags", [(3, 3, 3, 0), (3, 3, 9, 0), (2, 2, 2, 0), (4, 4, 6, 0), (5, 1, 12, 0),
                                                (3, 3, 5, 3), (3, 0, 4, 0), (0, 3, 1, 0), (1, 5, 3, 2)])
def test_random_models_vs_oracle(cw, tw, maxdict, tags):
    rng = np.random.default_rng(1000 * cw + 100 * tw + maxdict + tags)
    model, alpha = _random_model(rng, cw, tw, maxdict=maxdict, tags=tags)
    mb = encode_model(model)
    p, o = make(mb, tags=tags > 0), OraclePredictor(mb, predict_tags=tags > 0)
    sents = ["".join(rng.choice(list(alpha), size=rng.integers(1, 90))) for _ in range(700)]
    sents += ["".join(rng.choice(list(alpha), size=n)) for n in (1, 2, 31, 32, 33, 63, 64, 65, 96, 97, 300, 1025)]
    lens = [len(s.encode()) for s in sents]
    offs = np.zeros(len(lens) + 1, np.uint64)
    np.cumsum(lens, out=offs[1:])
    text = np.frombuffer("".join(sents).encode(), np.uint8)
    r = check_batch(p, o, text, offs, want_states=tags > 0)
    if tags:
        for i in (0, 5, len(sents) - 1):
            _, _, ocs, ots = o.predict(sents[i], states=True)
            c0 = int(r.char_offsets[i])
            if p.info["char_scorer"] == 2:
                assert r.char_states[c0:c0 + len(ocs)].tolist() == ocs.tolist()
            if p.info["type_scorer"] == 3:
                assert r.type_states[c0:c0 + len(ots)].tolist() == ots.tolist()


@pytest.mark.parametrize("cw,tw", [(3, 3), (3, 2), (4, 4), (0, 3), (6, 3)])  # (6, 3): general rows through k_tile_fast
def test_character_types_at_range_edges(cw, tw):
    """Both sides of every edge of CharacterType::get_type's ranges (sentence.rs:50-67), mixed into kana text: the tile
    kernels type BMP characters from a page table, and a page that holds two types needs its own sub-table -- U+4DBF | U+4DC0
    (the end of CJK Extension A inside page 4D) was the one that had none."""
    edges = [0x2F, 0x30, 0x39, 0x3A, 0x40, 0x41, 0x5A, 0x5B, 0x60, 0x61, 0x7A, 0x7B, 0x303F, 0x3040, 0x3096, 0x3097, 0x309F, 0x30A0,
             0x30FA, 0x30FB, 0x30FC, 0x30FF, 0x3100, 0x33FF, 0x3400, 0x4D00, 0x4DBF, 0x4DC0, 0x4DFF, 0x4E00, 0x9FFF, 0xA000,
             0xF8FF, 0xF900, 0xFAFF, 0xFB00, 0xFF0F, 0xFF10, 0xFF19, 0xFF1A, 0xFF20, 0xFF21, 0xFF3A, 0xFF3B, 0xFF40, 0xFF41, 0xFF5A,
             0xFF5B, 0xFF65, 0xFF66, 0xFF9F, 0xFFA0, 0x1FFFF, 0x20000, 0x2A6DF, 0x2A6E0, 0x2A6FF, 0x2A700, 0x2B73F, 0x2B740, 0x2B81F,
             0x2B820, 0x2CEAF, 0x2CEB0, 0x2F7FF, 0x2F800, 0x2FA1F, 0x2FA20,
             # and the edges of the UTF-8 encoding lengths, the surrogate gap and the code space
             0x01, 0x7F, 0x80, 0xE9, 0x3A9, 0x7FF, 0x800, 0xD7FF, 0xE000, 0xFFFD, 0xFFFF, 0x10000, 0x10FFFF]
    rng = np.random.default_rng(4242 + 10 * cw + tw)
    model, alpha = _random_model(rng, cw, tw, maxdict=3)
    # every type n-gram up to length 3 carries a weight: a wrong type changes a score
    tng = {}
    for n in (1, 2, 3):
        for k in range(6 ** n):
            g = bytes(1 + (k // 6 ** j) % 6 for j in range(n))
            tng[g] = rng.integers(-32767, 32768, size=max(2 * tw - n + 1, 0)).tolist()
    model["type_ngrams"] = list(tng.items()) if tw else []
    mb = encode_model(model)
    p, o = make(mb), OraclePredictor(mb)
    pool = [chr(c) for c in edges] + list(alpha)
    sents = ["".join(rng.choice(pool, size=rng.integers(1, 70))) for _ in range(600)]
    sents += ["".join(chr(c) for c in edges), "".join(chr(c) for c in range(0x4DB0, 0x4E10))]
    lens = [len(x.encode()) for x in sents]
    offs = np.zeros(len(lens) + 1, np.uint64)
    np.cumsum(lens, out=offs[1:])
    check_batch(p, o, np.frombuffer("".join(sents).encode(), np.uint8), offs)


@pytest.fixture(scope="module")
def synth_model():
    return synth.gen_model_bccwj_shaped(n_patterns=30000, sample_sentences=60000)


def test_synthetic_config2_shape(synth_model):
    # bccwj-suw-shaped (reduced pattern count so the CPU oracle builds in seconds): fast path, bit-exact
    p, o = make(synth_model), OraclePredictor(synth_model)
    assert p.info["fast_path"] == 1 and p.info["type_scorer"] == 2
    text, offs, _ = synth.gen_text(30000, 40)
    check_batch(p, o, text, offs)
    text, offs, _ = synth.gen_text(8000, ragged=True)
    check_batch(p, o, text, offs)


def test_fuzz_case_deep_key_vs_bigram():
    """A case the fuzzer found (tests/fuzz_gpu.py, seed 12345, model 1266): a (parent node, symbol) record whose parent id
    equals a code point has the low 42 key bits of a 2-character key; the first probe of that absent bigram landed on it."""
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fuzz_cases")
    mb = np.load(os.path.join(d, "case1266_model.npy")).tobytes()
    text, offs = np.load(os.path.join(d, "case1266_text.npy")), np.load(os.path.join(d, "case1266_offs.npy"))
    check_batch(make(mb), OraclePredictor(mb), text, offs)


def test_groups_that_do_not_fit_a_tile(synth_model):
    """64-sentence groups of long sentences (a group's text and slots exceed the tile buffers: slow path, totals published
    after the word-wise count) between groups of short ones (fast path), and mixed inside a group: offsets, scores and
    boundaries of every group depend on the totals of all groups before it."""
    p, o = make(synth_model), OraclePredictor(synth_model)
    short_t, short_o, _ = synth.gen_text(64 * 40, 40, seed=synth.TEXT_SEED + 31)
    long_t, long_o, _ = synth.gen_text(64 * 16, 170, seed=synth.TEXT_SEED + 32)
    mid_t, mid_o, _ = synth.gen_text(64 * 12, 61, seed=synth.TEXT_SEED + 33)   # 64 x 61 characters: just over the buffers
    sents = []
    for t, of in ((short_t, short_o), (long_t, long_o), (mid_t, mid_o)):
        sents.append([bytes(t[int(of[i]):int(of[i + 1])]) for i in range(len(of) - 1)])
    short_s, long_s, mid_s = sents
    order = []
    si = li = mi = 0
    for blk in range(52):
        kind = blk % 5
        if kind in (0, 2):
            order += short_s[si:si + 64]; si += 64
        elif kind == 1:
            order += long_s[li:li + 64]; li += 64
        elif kind == 3:
            order += mid_s[mi:mi + 64]; mi += 64
        else:   # mixed group
            order += short_s[si:si + 40] + long_s[li:li + 24]; si += 40; li += 24
    order = [x for x in order if x]
    enc = b"".join(order)
    offs = np.zeros(len(order) + 1, np.uint64)
    np.cumsum([len(x) for x in order], out=offs[1:])
    check_batch(p, o, np.frombuffer(enc, np.uint8), offs)


@pytest.mark.parametrize("budget", ["3", "300"])
def test_table_layout_variants_gpu(budget, monkeypatch):
    """dense 16-bit-seed tables (large dictionaries) and fat buckets, forced on a small model"""
    monkeypatch.setenv("VPT_SEED_BUDGET", budget)
    mb = synth.gen_model_bccwj_shaped(n_patterns=6000, sample_sentences=10000, dict_words=3000)
    p, o = make(mb), OraclePredictor(mb)
    text, offs, _ = synth.gen_text(3000, ragged=True)
    check_batch(p, o, text, offs)
    mb2 = synth.gen_model_bccwj_shaped(n_patterns=6000, sample_sentences=10000)
    p2, o2 = make(mb2), OraclePredictor(mb2)
    assert p2.info["fast_path"] == 1
    check_batch(p2, o2, text, offs)


def test_synthetic_config4_shape():
    # KyTea-shaped: + dictionary with words up to 16 chars (Variable-length rows in the reference; here inline
    # window + overflow rows on the deep records, tile kernel)
    mb = synth.gen_model_bccwj_shaped(n_patterns=20000, sample_sentences=40000, dict_words=20000)
    p, o = make(mb), OraclePredictor(mb)
    assert p.info["fast_path"] == 1 and p.info["max_char_pattern_len"] > 8
    text, offs, _ = synth.gen_text(10000, 40)
    check_batch(p, o, text, offs)
    text, offs, _ = synth.gen_text(3000, ragged=True)
    check_batch(p, o, text, offs)


def test_blob_roundtrip_and_device_api(synth_model):
    import ctypes as C
    import torch
    p = make(synth_model)
    blob = p.export_blob()
    q = vb.Predictor.from_blob(blob)
    text, offs, _ = synth.gen_text(5000, 40)
    a = p.predict_batch(text, offs)
    b = q.predict_batch(text, offs)
    assert np.array_equal(a.scores, b.scores) and np.array_equal(a.boundaries, b.boundaries)
    # device-pointer entry point
    dev = torch.device("cuda:0")
    n = len(offs) - 1
    d_text = torch.zeros(len(text) + 64, dtype=torch.uint8, device=dev)
    d_text[: len(text)] = torch.from_numpy(text.copy()).to(dev)
    d_off = torch.from_numpy(offs.astype(np.int64)).to(dev)
    ws = torch.empty(vb.lib().vpt_workspace_size(n), dtype=torch.uint8, device=dev)
    d_scores = torch.empty(len(text), dtype=torch.int32, device=dev)
    d_bounds = torch.empty(len(text), dtype=torch.uint8, device=dev)
    d_boff = torch.empty(n + 1, dtype=torch.int64, device=dev)
    d_status = torch.empty(n, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    rc = vb.lib().vpt_predict_batch_dev(q._h, d_text.data_ptr(), d_off.data_ptr(), n, ws.data_ptr(), ws.numel(),
                                        d_scores.data_ptr(), d_bounds.data_ptr(), d_boff.data_ptr(),
                                        d_status.data_ptr(), None, None, None, C.c_void_p(st))
    assert rc == 0, vb.lib().vpt_last_error()
    torch.cuda.synchronize()
    nb = int(d_boff[-1].item())
    assert nb == a.n_boundaries
    assert np.array_equal(d_scores[:nb].cpu().numpy(), a.scores)
    assert np.array_equal(d_bounds[:nb].cpu().numpy(), a.boundaries)
    assert int(d_status.abs().sum().item()) == 0


def test_many_tiny_sentences_offsets():
    """2.2 M two-character sentences: exercises the carry loop of the group scan (> 32768 groups), the
    chunked host path and tiny tiles.  Every sentence is the same text, so every score must be the same."""
    mb = read("model.bin")
    p, o = make(mb), OraclePredictor(mb)
    n = 2_200_000
    unit = "火星".encode()
    text = np.frombuffer(unit * n, np.uint8)
    offs = (np.arange(n + 1, dtype=np.uint64) * len(unit))
    r = p.predict_batch(text, offs)
    assert r.n_boundaries == n and np.array_equal(r.bound_offsets, np.arange(n + 1, dtype=np.uint64))
    want = o.predict("火星")[0]
    assert np.all(r.scores == want[0]) and int(r.status.sum()) == 0


def test_full_size_properties():
    """BASELINE config-2 size (1M x 40 chars) is checked through size-independent properties:
    (1) a batch equals the concatenation of its halves (sentences are independent);
    (2) scoring is invariant to the sentence's position in the batch (permutation);
    (3) the sample scored by the oracle matches bit-exactly."""
    mb = synth.gen_model_bccwj_shaped(n_patterns=30000, sample_sentences=60000)
    p, o = make(mb), OraclePredictor(mb)
    n = 1_000_000
    text, offs, _ = synth.gen_text(n, 40)
    full = p.predict_batch(text, offs)
    assert full.n_boundaries == n * 39 and int(full.status.sum()) == 0
    h = n // 2
    a = p.predict_batch(text, offs[: h + 1])
    b = p.predict_batch(text, offs[h:])
    assert np.array_equal(np.concatenate([a.scores, b.scores]), full.scores)
    assert np.array_equal(np.concatenate([a.boundaries, b.boundaries]), full.boundaries)
    assert np.array_equal(full.boundaries, (full.scores > 0).astype(np.uint8))
    # checksum of a strided sample against the oracle
    idx = np.arange(0, n, 997)
    for i in idx[:300]:
        s = bytes(text[int(offs[i]):int(offs[i + 1])]).decode()
        assert full.sentence_scores(i).tolist() == o.predict(s)[0].tolist()
