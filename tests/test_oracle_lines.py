"""The oracle's restatement of the reference CLI loop (`predict --no-norm`, predict/src/main.rs:126-150) against
the per-sentence oracle (itself pinned to the reference's vectors) and Rust's `BufRead::lines` rules as Python's
own line handling states them."""
import json
import os

import vaporetto_b200 as vb
from vpt_testlib.oracle import OraclePredictor, lib as oracle_lib

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rust_lines(data: bytes):
    """std::io::BufRead::lines: split at '\\n', drop one '\\r' before it, no empty line after a trailing '\\n'."""
    if not data:
        return []
    parts = data.split(b"\n")
    terminated = parts[:-1]
    out = [p[:-1] if p.endswith(b"\r") else p for p in terminated]
    if parts[-1]:
        out.append(parts[-1])  # unterminated last line keeps a trailing '\r'
    return out


def fullwidth_map():
    with open(os.path.join(GOLDEN, "kytea_fullwidth_map.json")) as f:
        return {int(k): v for k, v in json.load(f).items()}


def escape(tok: str) -> str:
    return tok.replace("\\", "\\\\").replace(" ", "\\ ").replace("/", "\\/")


def char_type(c: str) -> int:
    return int(vb.Sentence.from_raw(c).char_types()[0])


def expected(o, data: bytes, no_norm: bool = True, wsconst: str = "") -> bytes:
    fw = fullwidth_map()
    out = []
    for line in rust_lines(data):
        try:
            s = line.decode("utf-8")
            ok = len(s) > 0 and "\x00" not in s
        except UnicodeDecodeError:
            ok = False
        if not ok:
            out.append(b"")
        elif no_norm and not wsconst:
            out.append(o.tokenize(s).encode())
        else:
            # predict on the filtered line, put its boundaries on the original line (predict/src/main.rs:154-166);
            # --wsconst filters run on the predicted sentence (kytea_wsconst.rs:27-44)
            pre = s if no_norm else "".join(chr(fw.get(ord(c), ord(c))) for c in s)
            _, bounds = o.predict(pre)
            bounds = bounds.tolist()
            types = [char_type(c) for c in pre]
            for letter in wsconst:
                t = "DRHTKO".index(letter) + 1
                for i in range(len(types) - 1):
                    if types[i] == t and types[i + 1] == t:
                        bounds[i] = 0
            toks, start = [], 0
            for i, b in enumerate(bounds):
                if b == 1:
                    toks.append(s[start:i + 1])
                    start = i + 1
            toks.append(s[start:])
            out.append(" ".join(escape(t) for t in toks).encode())
    return b"".join(x + b"\n" for x in out)


def test_fullwidth_map_matches_reference_fixture():
    """Every code point up to U+FFFF (and a few beyond): the oracle's table and the library's arithmetic form
    (csrc/textnorm.hpp, the function the kernels apply) against the map extracted from the reference source."""
    fw = fullwidth_map()
    assert len(fw) == 96
    L, O = vb.lib(), oracle_lib()
    for c in list(range(0, 0x10000)) + [0x10000, 0x1F600, 0x2A6DF, 0x10FFFF]:
        want = fw.get(c, c)
        assert O.ora_kytea_fullwidth(c) == want, hex(c)
        assert L.vpt_kytea_fullwidth(c) == want, hex(c)


def test_cli_loop_matches_per_line_oracle():
    with open(os.path.join(GOLDEN, "model.bin"), "rb") as f:
        o = OraclePredictor(f.read())
    cases = [
        b"", b"\n", b"\n\n", b"\r\n", b"\r", b"a", b"a\n", b"a\r\n", b"a\r\r\n", b"a\rb\n",
        "まぁ社長は火星猫だ".encode(),
        "まぁ社長は火星猫だ\r\nまぁ社長は火星猫だ".encode(),
        "火星 猫/です\\ね\n\n 火星\n/\n\\\n".encode(),
        "まぁ\x00社長\n火星猫\n".encode(),
        b"\xe3\x81\n" + "火星猫\n".encode() + b"\xff\n",
    ]
    cases += [
        "Vaporetto is a tokenizer. (v0.6.5) - 100% [test]\n".encode(),
        "ｶﾞｰﾃﾞﾝ－ハウス―A–B─C ｢x｣ ～ ､ ･ ｡\n".encode(),
        "a/b c\\d 1.5 -3 \"q\" 'r' #$;^`|~\n".encode(),
    ]
    for data in cases:
        for no_norm in (True, False):
            for ws in ("", "D", "RK", "DRHTKO"):
                got, nl = o.tokenize_lines(data, no_norm=no_norm, wsconst=ws)
                assert nl == len(rust_lines(data))
                assert got == expected(o, data, no_norm, ws), (data, no_norm, ws)


def test_wsconst_reference_vectors():
    """kytea_wsconst.rs:57-80: the filter on given boundaries ("5 00 0" -> "5000", "20 21 年 8 月 2 4 日" ->
    "2021 年 8 月 24 日"), through a model that cuts everywhere (bias > 0, no features)."""
    from vpt_testlib.bincode_model import encode_model
    o = OraclePredictor(encode_model(dict(char_ngrams=[], type_ngrams=[], dict=[], bias=1, char_window=1, type_window=1,
                                          tag_models=[])))
    got, _ = o.tokenize_lines("5\n5000\n2021年8月24日\n".encode(), no_norm=True, wsconst="D")
    assert got.decode() == "5\n5000\n2021 年 8 月 24 日\n"
    got, _ = o.tokenize_lines("2021年8月24日\n".encode(), no_norm=True)
    assert got.decode() == "2 0 2 1 年 8 月 2 4 日\n"


def test_docs_tok_through_cli_loop():
    """The reference's documented tokenisation (tests/golden/docs.tok) through the whole-buffer loop."""
    with open(os.path.join(GOLDEN, "model.bin"), "rb") as f:
        o = OraclePredictor(f.read())
    got, nl = o.tokenize_lines("まぁ社長は火星猫だ\nまぁ社長は火星猫だ\n".encode(), no_norm=True)
    assert nl == 2
    assert got.decode() == "まぁ 社長 は 火星 猫 だ\n" * 2


def test_tantivy_pipeline_vectors():
    """vaporetto_tantivy's token_stream (lib.rs:160-199: pre-filter -> predict -> wsconst post-filters -> tokens of the
    original text) is the same pipeline as the CLI loop: its unit tests pin the oracle's restatement."""
    from golden import reference_kat as kat
    with open(os.path.join(GOLDEN, "tantivy_model.bin"), "rb") as f:
        o = OraclePredictor(f.read())
    for text, ws, want in kat.TANTIVY_PIPELINE:
        got, nl = o.tokenize_lines((text + "\n").encode(), no_norm=False, wsconst=ws)
        assert nl == 1 and got.decode() == want + "\n", (text, ws)
    assert o.tokenize_lines(b"", no_norm=False) == (b"", 0)          # lib.rs:255-260: no tokens for ""


def test_tokenized_escape_vector():
    """sentence.rs:2695-2701 through the oracle (a model built to give the annotated boundaries) and through the
    library's host-side `Sentence::write_tokenized_text` with the boundaries set by hand."""
    from golden import reference_kat as kat
    from vpt_testlib.bincode_model import encode_model
    case = kat.TOKENIZED_ESCAPE
    o = OraclePredictor(encode_model(case["model"]))
    _, bounds = o.predict(case["text"])
    assert bounds.tolist() == case["boundaries"]
    assert o.tokenize(case["text"]) == case["tokenized"]
    got, _ = o.tokenize_lines((case["text"] + "\n").encode(), no_norm=True)
    assert got.decode() == case["tokenized"] + "\n"
    s = vb.Sentence.from_raw(case["text"])
    s.boundaries_mut()[:] = case["boundaries"]
    assert s.write_tokenized_text() == case["tokenized"]
