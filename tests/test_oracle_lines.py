"""The oracle's restatement of the reference CLI loop (`predict --no-norm`, predict/src/main.rs:126-150) against
the per-sentence oracle (itself pinned to the reference's vectors) and Rust's `BufRead::lines` rules as Python's
own line handling states them."""
import json
import os

import vaporetto_b200 as vb
from vpt_testlib.oracle import OraclePredictor, lib as oracle_lib

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


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
    """Every code point: the oracle's table and the library's arithmetic form
    (csrc/textnorm.hpp, the function the kernels apply) against the map extracted from the reference source."""
    fw = fullwidth_map()
    assert len(fw) == 96
    L, O = vb.lib(), oracle_lib()
    for c in range(0, 0x110000):
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


# ---- --wsconst G: ConcatGraphemeClustersFilter (concat_grapheme_clusters.rs:10-35) ---------------------------------------

GRAPHEME_POOL = (
    "aZ 0。あア漢ｶﾞ" "\r\n\t‍‌゙゚̀́️︎⃣"
    "\U0001f468\U0001f469\U0001f466\U0001f44f\U0001f3fd\U0001f3fb❤©™\U0001f1ef\U0001f1f5\U0001f1fa\U0001f1f8\U0001f1e9"
    "각가각ힰퟋ؀؅ःाक्ष্ી്കำຳ"
    "ཀཱါ᭄ꢴ\U000110bd\U00011000\U0001d165\U000e0020\U000e007f\U000e0001 ­"
)


def _regex_cluster_lengths(text):
    import regex
    return [len(m) for m in regex.findall(r"\X", text)]


def test_grapheme_rules_against_regex():
    """The oracle's rule engine (UAX #29 extended grapheme clusters, what unicode-segmentation's graphemes(true) yields)
    against the `regex` module's \\X on 30000 random strings over marks, emoji sequences, flags, Hangul jamo, Indic
    conjuncts, prepended marks and controls."""
    import random
    from vpt_testlib.oracle import grapheme_lengths
    rng = random.Random(29)
    for it in range(30000):
        n = rng.randint(1, 14)
        text = "".join(rng.choice(GRAPHEME_POOL) for _ in range(n))
        assert grapheme_lengths(text) == _regex_cluster_lengths(text), [hex(ord(c)) for c in text]
    # every code point that starts a range of a non-default class, next to a few neighbours
    import re
    src = open(os.path.join(HERE, "..", "oracle", "grapheme_tables.hpp")).read()
    firsts = [int(m.group(1), 16) for m in re.finditer(r"\{0x([0-9A-F]+), 0x", src)]
    for c in firsts:
        if 0xD800 <= c <= 0xDFFF or c == 0:
            continue
        for text in ("あ" + chr(c) + "あ", chr(c) + chr(c), "क" + chr(c) + "क",
                     "\U0001f468" + chr(c) + "\U0001f469"):
            assert grapheme_lengths(text) == _regex_cluster_lengths(text), [hex(ord(x)) for x in text]


def test_grapheme_filter_reference_vectors():
    """concat_grapheme_clusters.rs:36-83 (the filter's unit tests): clusters written as one token each."""
    from vpt_testlib.bincode_model import encode_model
    from golden import reference_kat as kat
    # a model that splits everywhere (positive bias, no patterns): the filter alone decides
    m = {**kat.TOKENIZED_ESCAPE["model"], "char_ngrams": [], "type_ngrams": [], "dict": [], "bias": 1, "tag_models": []}
    o = OraclePredictor(encode_model(m))
    for text, want in (
        ("‍", "‍"),
        ("\U0001f468‍\U0001f469‍\U0001f466", "\U0001f468‍\U0001f469‍\U0001f466"),
        ("\U0001f44f\U0001f3fd", "\U0001f44f\U0001f3fd"),
        ("これは手\U0001f44f\U0001f3fdです", "こ れ は 手 \U0001f44f\U0001f3fd で す"),
    ):
        got, _ = o.tokenize_lines((text + "\n").encode(), no_norm=True, wsconst="G")
        assert got.decode() == want + "\n", text
        assert o.tokenize_lines((text + "\n").encode(), no_norm=True)[0].decode() == " ".join(text) + "\n"


def test_host_grapheme_filter_matches_oracle():
    """vpt_concat_grapheme_clusters (the rule engine the device kernel runs, as a per-character state machine) against
    the oracle's look-back restatement on random strings, and the reference's vectors through the Sentence mirror."""
    import random
    from vpt_testlib.oracle import grapheme_lengths
    rng = random.Random(31)
    for it in range(20000):
        text = "".join(rng.choice(GRAPHEME_POOL.replace("\0", "")) for _ in range(rng.randint(1, 40)))
        s = vb.Sentence.from_raw(text)
        s.boundaries_mut()[:] = 1
        s.concat_grapheme_clusters()
        lens, run = [], 1
        for b in s.boundaries().tolist():
            if b:
                lens.append(run)
                run = 1
            else:
                run += 1
        lens.append(run)
        assert lens == grapheme_lengths(text), [hex(ord(c)) for c in text]
    s = vb.Sentence.from_raw("これは手\U0001f44f\U0001f3fdです")
    s.boundaries_mut()[:] = 1
    s.concat_grapheme_clusters()
    assert s.write_tokenized_text() == "こ れ は 手 \U0001f44f\U0001f3fd で す"


def test_split_linebreaks_reference_vectors():
    """split_linebreaks.rs:45-76 (the filter's unit tests) through the Sentence mirror (vpt_split_linebreaks)."""
    for text, want in (("前の行\n次の行", "前の行 \n 次の行"), ("前の行\r次の行", "前の行 \r 次の行"),
                       ("前の行\r\n次の行", "前の行 \r \n 次の行")):
        s = vb.Sentence.from_raw(text)
        s.boundaries_mut()[:] = 0       # Sentence::from_tokenized of an unsegmented string: no boundary anywhere
        s.split_linebreaks()
        assert s.write_tokenized_text() == want
    s = vb.Sentence.from_raw("\n")
    s.split_linebreaks()
    assert s.boundaries().size == 0


def test_grapheme_classes_at_range_ends():
    """Both ends of every range of the grapheme property table (and the code points just outside): the oracle's engine
    against regex's \\X, and the library's per-character state machine (the engine the device kernel runs) against the
    oracle, in five contexts (kana, doubled, Devanagari consonants, emoji, Hangul jamo).  A full sweep of the code space
    (every code point, the same contexts) was run once for both: 0 mismatches."""
    import re
    from vpt_testlib.oracle import grapheme_lengths
    src = open(os.path.join(HERE, "..", "oracle", "grapheme_tables.hpp")).read()
    cps = set()
    for m in re.finditer(r"\{0x([0-9A-F]+), 0x([0-9A-F]+)", src):
        lo, hi = int(m.group(1), 16), int(m.group(2), 16)
        cps.update((lo - 1, lo, hi, hi + 1))
    ctxs = [lambda ch: "あ" + ch + "あ", lambda ch: ch + ch, lambda ch: "क" + ch + "क",
            lambda ch: "\U0001f468" + ch + "\U0001f469", lambda ch: "ᄀ" + ch + "ᅡ"]

    def product(text):
        s = vb.Sentence.from_raw(text)
        s.boundaries_mut()[:] = 1
        s.concat_grapheme_clusters()
        lens, run = [], 1
        for b in s.boundaries().tolist():
            if b:
                lens.append(run)
                run = 1
            else:
                run += 1
        return lens + [run]

    for c in sorted(cps):
        if c < 1 or c > 0x10FFFF or 0xD800 <= c <= 0xDFFF:
            continue
        for f in ctxs:
            text = f(chr(c))
            want = _regex_cluster_lengths(text)
            assert grapheme_lengths(text) == want, [hex(ord(x)) for x in text]
            assert product(text) == want, [hex(ord(x)) for x in text]
