"""GPU parity tests of vpt_tokenize_lines (device-side line splitting + tokenised output)
against the oracle's restatement of the reference CLI loop (predict/src/main.rs:126-150).  Byte-exact."""
import os

import numpy as np
import pytest

import vaporetto_b200 as vb
from vpt_testlib import synth
from vpt_testlib.bincode_model import encode_model
from golden import reference_kat as kat
from vpt_testlib.oracle import OraclePredictor
from test_gpu_parity import _random_model, make, read

pytestmark = pytest.mark.gpu


def check(p, o, data: bytes, wsconsts=("", "D", "HK", "DRHTKO", "G", "GD")):
    for no_norm in (True, False):
        for ws in wsconsts:
            got, nl = p.tokenize_lines(data, no_norm=no_norm, wsconst=ws)
            want, wl = o.tokenize_lines(data, no_norm=no_norm, wsconst=ws)
            assert nl == wl
            assert got.tobytes() == want, (no_norm, ws, data[:200])


@pytest.fixture(scope="module")
def kat_pair():
    mb = read("model.bin")
    return make(mb), OraclePredictor(mb)


def test_lines_semantics(kat_pair, monkeypatch):
    p, o = kat_pair
    cases = [
        b"",
        b"\n",
        b"\n\n\n",
        b"\r\n",
        b"\r",
        b"a",
        b"a\n",
        b"a\r\n",
        b"a\r\r\n",
        b"a\rb\n",
        "まぁ社長は火星猫だ".encode(),
        "まぁ社長は火星猫だ\n".encode(),
        "まぁ社長は火星猫だ\r\nまぁ社長は火星猫だ".encode(),
        "火星 猫/です\\ね\n\n 火星\n/\n\\\n".encode(),                 # escapes of ' ', '/', '\\'
        "まぁ\x00社長\n火星猫\n".encode(),                                # NUL line -> empty line
        b"\xe3\x81\n" + "火星猫\n".encode() + b"\xff\xfe\n\xc0\x80\n",      # invalid UTF-8 lines -> empty lines
        ("火星猫だ" * 3000 + "\n" + "まぁ社長は" * 7 + "\n").encode(),     # a line larger than a tile
        "\n".join("まぁ社長は火星猫だ"[: 1 + i % 9] for i in range(500)).encode(),
        "\U00020000\U0002a6df火星été ab12 ｶﾀｶﾅ\n".encode(),       # 4-, 2-byte characters
        "Vaporetto is a tokenizer. (v0.6.5) - 100% [test]\n".encode(),                 # KyteaFullwidthFilter sources
        "ｶﾞｰﾃﾞﾝ－ハウス―A–B─C ｢x｣ ～ ､ ･ ｡\n".encode(),
        "a/b c\\d 1.5 -3 \"q\" 'r' #$;^`|~\n".encode(),
    ]
    for chunk in ("", "64", "200"):
        if chunk:
            monkeypatch.setenv("VPT_CHUNK_BYTES", chunk)
        for data in cases:
            check(p, o, data)


def test_wsconst_reference_vectors():
    # kytea_wsconst.rs:57-80 through a model that cuts everywhere (bias > 0, no features)
    mb = encode_model(dict(char_ngrams=[], type_ngrams=[], dict=[], bias=1, char_window=1, type_window=1, tag_models=[]))
    p = make(mb)
    got, _ = p.tokenize_lines("5\n5000\n2021年8月24日\n".encode(), no_norm=True, wsconst="D")
    assert got.tobytes().decode() == "5\n5000\n2021 年 8 月 24 日\n"
    with pytest.raises(vb.VaporettoError):
        p.tokenize_lines(b"a\n", wsconst="X")


def test_grapheme_filter_on_device():
    """--wsconst G (ConcatGraphemeClustersFilter, concat_grapheme_clusters.rs:10-35): the reference's unit-test vectors and
    random lines over marks, emoji sequences, flags, jamo, conjuncts and controls, device vs oracle, byte-exact."""
    import random
    from test_oracle_lines import GRAPHEME_POOL
    mb = encode_model(dict(char_ngrams=[], type_ngrams=[], dict=[], bias=1, char_window=1, type_window=1, tag_models=[]))
    p, o = make(mb), OraclePredictor(mb)
    got, _ = p.tokenize_lines("\u200d\n\U0001f468\u200d\U0001f469\u200d\U0001f466\n\U0001f44f\U0001f3fd\nこれは手\U0001f44f\U0001f3fdです\n".encode(),
                              no_norm=True, wsconst="G")
    assert got.tobytes().decode() == ("\u200d\n\U0001f468\u200d\U0001f469\u200d\U0001f466\n\U0001f44f\U0001f3fd\n"
                                      "こ れ は 手 \U0001f44f\U0001f3fd で す\n")
    rng = random.Random(5)
    pool = GRAPHEME_POOL.replace("\n", "").replace("\r", "")
    lines = []
    for _ in range(20000):
        n = rng.choice((1, 2, 3, 5, 8, 13, 31, 32, 33, 40, 64, 65, 130, 300))
        if rng.random() < 0.5:   # mostly plain text with a few special characters (the windows' fast path and its edges)
            line = "".join(rng.choice(pool) if rng.random() < 0.05 else rng.choice("あいう漢字カナab1 ") for _ in range(n))
        else:
            line = "".join(rng.choice(pool) for _ in range(n))
        lines.append(line)
    data = ("\n".join(lines) + "\n").encode()
    for no_norm in (True, False):
        for ws in ("G", "GHK"):
            got, nl = p.tokenize_lines(data, no_norm=no_norm, wsconst=ws)
            want, wl = o.tokenize_lines(data, no_norm=no_norm, wsconst=ws)
            assert nl == wl == len(lines)
            assert got.tobytes() == want, (no_norm, ws)


def test_tantivy_pipeline_vectors():
    # vaporetto_tantivy/src/lib.rs:160-199 + its tests (:263-399): pre-filter -> predict -> wsconst -> original text
    p = make(read("tantivy_model.bin"))
    for text, ws, want in kat.TANTIVY_PIPELINE:
        got, nl = p.tokenize_lines((text + "\n").encode(), wsconst=ws)
        assert nl == 1 and got.tobytes().decode() == want + "\n", (text, ws)


def test_tokenized_escape_vector():
    # sentence.rs:2695-2701 on the device: boundaries from a model built to give the annotation, escapes of ' ' and '\\'
    case = kat.TOKENIZED_ESCAPE
    p = make(encode_model(case["model"]))
    s = vb.Sentence.from_raw(case["text"])
    p.predict(s)
    assert s.boundaries().tolist() == case["boundaries"]
    got, _ = p.tokenize_lines((case["text"] + "\n").encode(), no_norm=True)
    assert got.tobytes().decode() == case["tokenized"] + "\n"


def test_lines_out_capacity(kat_pair):
    p, o = kat_pair
    data = "まぁ社長は火星猫だ\n".encode() * 10
    want, _ = o.tokenize_lines(data, no_norm=True)
    out = np.empty(len(want), np.uint8)
    got, nl = p.tokenize_lines(data, out=out, no_norm=True)
    assert got.tobytes() == want and nl == 10
    with pytest.raises(vb.VaporettoError):
        p.tokenize_lines(data, out=np.empty(len(want) - 1, np.uint8))


def _random_lines(rng, n_lines, alphabet, maxlen):
    parts = []
    for _ in range(n_lines):
        r = rng.random()
        if r < 0.05:
            line = b""
        elif r < 0.08:
            line = bytes(rng.integers(0, 256, rng.integers(1, 12)).astype(np.uint8))  # mostly malformed
        else:
            n = int(rng.integers(1, maxlen))
            line = "".join(alphabet[int(i)] for i in rng.integers(0, len(alphabet), n)).encode()
        line = line.replace(b"\n", b"")
        parts.append(line + (b"\r\n" if rng.random() < 0.2 else b"\n"))
    data = b"".join(parts)
    if rng.random() < 0.5:
        data = data[:-1] if not data.endswith(b"\r\n") else data[:-2]
    return data


@pytest.mark.parametrize("cw,tw,maxdict", [(3, 3, 3), (3, 3, 9), (2, 4, 6)])
def test_random_lines_vs_oracle(cw, tw, maxdict, monkeypatch):
    rng = np.random.default_rng(cw * 100 + tw * 10 + maxdict)
    for it in range(6):
        m, alpha = _random_model(rng, cw, tw, maxdict=maxdict)
        alphabet = list(alpha) + list(" /\\é\U00020000") + list("ab.-ｱ－―｢､")
        mb = encode_model(m)
        p, o = make(mb), OraclePredictor(mb)
        monkeypatch.setenv("VPT_CHUNK_BYTES", str(int(rng.choice([64, 1000, 1 << 16, 8 << 20]))))
        check(p, o, _random_lines(rng, int(rng.integers(1, 400)), alphabet, 80))
        check(p, o, _random_lines(rng, 5, alphabet, 3000))


def test_lines_full_size():
    """BASELINE config-2 shape (300 K-pattern model, 200 K lines): byte-exact against the oracle, and the
    size-independent properties: removing the inserted bytes gives the input back; line count is kept."""
    mb = synth.gen_model_bccwj_shaped(n_patterns=300_000, sample_sentences=200_000)
    p, o = make(mb), OraclePredictor(mb)
    text, offs, _ = synth.gen_text(200_000, 40, seed=99)
    lines = [text[int(offs[i]):int(offs[i + 1])].tobytes() for i in range(len(offs) - 1)]
    data = b"\n".join(lines) + b"\n"
    got, nl = p.tokenize_lines(data, no_norm=True)
    assert nl == len(lines)
    want, _ = o.tokenize_lines(data, no_norm=True)
    assert got.tobytes() == want
    got_n, _ = p.tokenize_lines(data, wsconst="DK")        # CLI default (full-width normalisation) + --wsconst D K
    want_n, _ = o.tokenize_lines(data, wsconst="DK")
    assert got_n.tobytes() == want_n
    g = got.tobytes()
    assert g.count(b"\n") == len(lines)
    # the synthetic text has ' ' but no '/', '\\' or NUL: undoing the escapes and dropping the separators
    # restores the input
    assert b"/" not in data and b"\\" not in data
    assert g.replace(b"\\ ", b"\x00").replace(b" ", b"").replace(b"\x00", b" ") == data


def test_lines_with_tags_reference_model():
    """--predict-tags through vpt_tokenize_lines_tags: the bundled model's documented output (README / predictor.rs doctest:
    "まぁ/名詞/マー 社長/名詞/シャチョー ...") and mixed lines, with and without the full-width pre-filter and post-filters,
    byte-exact against the oracle's CLI loop (fill_tags on the predicted sentence, main.rs:157-166)."""
    mb = read("model.bin")
    p, o = vb.Predictor(vb.Model.read(mb), predict_tags=True), OraclePredictor(mb, predict_tags=True)
    data = ("まぁ社長は火星猫だ\r\n\nまぁ良いだろう\nVaporetto 1.5/2 a\\b\n社長は社長だ火星猫\n" + "火星猫は社長だ" * 40 + "\n" + "猫\n").encode()
    got, nl = p.tokenize_lines(data, no_norm=True, predict_tags=True)
    assert got.tobytes().decode().split("\n")[0] == "まぁ/名詞/マー 社長/名詞/シャチョー は/助詞/ワ 火星/名詞/カセー 猫/名詞/ネコ だ/助動詞/ダ"
    for no_norm in (True, False):
        for ws in ("", "K", "GD"):
            got, nl = p.tokenize_lines(data, no_norm=no_norm, wsconst=ws, predict_tags=True)
            want, wl = o.tokenize_lines(data, no_norm=no_norm, wsconst=ws, predict_tags=True)
            assert nl == wl and got.tobytes() == want, (no_norm, ws)
    p0 = make(mb)
    with pytest.raises(vb.VaporettoError):
        p0.tokenize_lines(data, predict_tags=True)


def test_lines_with_tags_synthetic(monkeypatch):
    """1 500 tag models (tokens of ASCII and Japanese characters, tag strings with characters that need escapes) on a
    30 000-pattern model, 20 000 lines, default normalisation (tokens are looked up by their full-width image)."""
    from vpt_testlib import synth
    mb = synth.gen_model_bccwj_shaped(n_patterns=30_000, sample_sentences=50_000, tag_models=1_500)
    p, o = vb.Predictor(vb.Model.read(mb), predict_tags=True), OraclePredictor(mb, predict_tags=True)
    text, offs, _ = synth.gen_text(20_000, 40, seed=synth.TEXT_SEED + 21)
    lines = [text[int(offs[i]):int(offs[i + 1])].tobytes() for i in range(len(offs) - 1)]
    lines[3] = b""
    lines[17] = b"a\x00b"
    lines[18] = b"\xff\xfe"
    data = b"\n".join(lines) + b"\n"
    monkeypatch.setenv("VPT_CHUNK_BYTES", "300000")
    for no_norm in (False, True):
        got, nl = p.tokenize_lines(data, no_norm=no_norm, predict_tags=True)
        want, wl = o.tokenize_lines(data, no_norm=no_norm, predict_tags=True)
        assert nl == wl == len(lines)
        assert got.tobytes() == want, no_norm
    assert got.tobytes().count(b"/") > 1000
