"""CPU checks of the product's host-side builder (merged rows, reversed-suffix node table, perfect hash,
flat blob) by interpreting the blob with the kernels' probe sequence (tests/native/host_emul.cpp) and
comparing with the oracle.  No GPU, no product compute path involved."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from golden import reference_kat as kat
from vpt_testlib.bincode_model import encode_model
from vpt_testlib.oracle import OraclePredictor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden")
CSRC = os.path.join(ROOT, "vaporetto_b200", "csrc")


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(HERE, "native", "libhost_emul.so")
    srcs = [os.path.join(HERE, "native", "host_emul.cpp")] + [os.path.join(CSRC, f) for f in
                                                             ("predictor_build.cpp", "builder.cpp", "model.cpp", "tags_build.cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in ("builder.hpp", "keys.hpp", "predictor_build.hpp", "common.hpp", "tags.hpp", "tags_token.hpp",
                                                     "textnorm.hpp")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in deps):
        # (tags.hpp declares the launch interface next to the tables: cuda_runtime.h for the types only, nothing is linked)
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I" + cuda_inc, "-o", so] + srcs)
    L = C.CDLL(so)
    L.emul_predict.restype = C.c_long
    L.emul_predict.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p]
    L.emul_last_error.restype = C.c_char_p
    L.emul_predict_tags.restype = C.c_long
    L.emul_predict_tags.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return L


def run(L, model_bytes, text, tags=False):
    b = text.encode()
    sc = np.zeros(len(b) + 1, np.int32)
    cs = np.zeros(len(b) + 1, np.uint32)
    ts = np.zeros(len(b) + 1, np.uint32)
    info = np.zeros(4, np.int32)
    n = L.emul_predict(model_bytes, len(model_bytes), int(tags), b, len(b), sc.ctypes.data, cs.ctypes.data,
                       ts.ctypes.data, info.ctypes.data)
    assert n > 0, L.emul_last_error()
    return sc[: n - 1].tolist(), cs[:n].tolist(), ts[:n].tolist(), info.tolist()


@pytest.mark.parametrize("name", sorted(kat.SCORE_CASES))
def test_tables_on_reference_vectors(emul, name):
    case = kat.SCORE_CASES[name]
    sc, _, _, info = run(emul, encode_model(case["model"]), case["text"])
    assert sc == case["scores"]


@pytest.mark.parametrize("name", sorted(kat.TAG_SCORE_CASES))
def test_tables_on_reference_vectors_tags(emul, name):
    case = kat.TAG_SCORE_CASES[name]
    mb = encode_model(case["model"])
    sc, cs, ts, info = run(emul, mb, case["text"], tags=True)
    assert sc == case["scores"]
    o = OraclePredictor(mb, predict_tags=True)
    _, _, ocs, ots = o.predict(case["text"], states=True)
    assert cs == ocs.tolist()
    assert ts == ots.tolist()


def test_fast_path_selected(emul):
    # n-gram-only W=3 models and short dictionary words fit the 6-wide inline rows
    _, _, _, info = run(emul, encode_model(kat.PREDICTOR_TEST_MODEL), "この人は地球人だ")
    assert info[0] == 1
    _, _, _, info = run(emul, encode_model(kat.CHAR_ADD_SCORES_3["model"]), "我らは全世界の国民")
    assert info[0] == 1  # 5-char dictionary words: inline window + overflow rows on the deep records
    wide = dict(char_ngrams=[("界", list(range(1, 11)))], bias=0, char_window=5, type_window=0)  # 10-wide n-gram row
    _, _, _, info = run(emul, encode_model(wide), "我らは全世界の国民")
    assert info[0] == 0
    _, _, _, info = run(emul, encode_model(kat.TYPE_ADD_SCORES["model"]), "我らは全世界の国民")
    assert info[0] == 0 and info[2] == 1  # type window 4: automaton variant


def test_fixture_models(emul):
    for fn, texts in (("model.bin", list(kat.MODEL_BIN_SCORES)), ("tantivy_model.bin", [t for t, _ in kat.TANTIVY_TOKENIZE])):
        with open(os.path.join(GOLDEN, fn), "rb") as f:
            mb = f.read()
        for tags in (False, True):
            o = OraclePredictor(mb, predict_tags=tags)
            for text in texts:
                sc, cs, ts, _ = run(emul, mb, text, tags=tags)
                osc, _, ocs, ots = o.predict(text, states=True)
                assert sc == osc.tolist()
                if tags:
                    assert cs == ocs.tolist() and ts == ots.tolist()


ALPHA = "あいうアイ人火星地球aB1。"
chars = st.sampled_from(list(ALPHA))
w = st.integers(-40000, 40000)


@st.composite
def models(draw):
    cw = draw(st.integers(0, 5))
    tw = draw(st.integers(0, 5))
    cng = {}
    for _ in range(draw(st.integers(0, 8))):
        g = "".join(draw(st.lists(chars, min_size=1, max_size=5)))
        full = max(2 * cw - len(g) + 1, 0)
        cng[g] = draw(st.lists(w, min_size=0, max_size=full + 1))
    dic = []
    for _ in range(draw(st.integers(0, 6))):
        g = "".join(draw(st.lists(chars, min_size=1, max_size=9)))
        dic.append((g, draw(st.lists(w, min_size=0, max_size=len(g) + 2)), ""))
    tng = {}
    for _ in range(draw(st.integers(0, 6))):
        g = bytes(draw(st.lists(st.integers(1, 6), min_size=1, max_size=5)))
        full = max(2 * tw - len(g) + 1, 0)
        tng[g] = draw(st.lists(w, min_size=0, max_size=full + 1))
    tms = []
    for t in range(draw(st.integers(0, 2))):
        cn = [("".join(draw(st.lists(chars, min_size=1, max_size=4))),
               [(draw(st.integers(0, cw)), draw(st.lists(w, min_size=1, max_size=3)))]) for _ in range(draw(st.integers(0, 3)))]
        tn = [(bytes(draw(st.lists(st.integers(1, 6), min_size=1, max_size=4))),
               [(draw(st.integers(0, tw)), draw(st.lists(w, min_size=1, max_size=3)))]) for _ in range(draw(st.integers(0, 3)))]
        tms.append(dict(token="tok%d" % t, tags=[["x", "y"]], char_ngrams=cn, type_ngrams=tn, bias=[1, 2]))
    return dict(char_ngrams=list(cng.items()), type_ngrams=list(tng.items()), dict=dic,
                bias=draw(st.sampled_from([0, 5, -7, 2**31 - 1])), char_window=cw, type_window=tw, tag_models=tms)


@settings(max_examples=300, deadline=None)
@given(models(), st.lists(chars, min_size=1, max_size=24), st.booleans())
def test_tables_equal_oracle(emul, model, text, tags):
    text = "".join(text)
    mb = encode_model(model)
    o = OraclePredictor(mb, predict_tags=tags)
    osc, _, ocs, ots = o.predict(text, states=True)
    sc, cs, ts, info = run(emul, mb, text, tags=tags)
    assert sc == osc.tolist()
    if tags and model["tag_models"]:
        if o.type_variant == 1:
            assert ts == ots.tolist()
        if model["char_window"] and (model["char_ngrams"] or model["dict"]):
            assert cs == ocs.tolist()


@pytest.mark.parametrize("budget", ["3", "40", "300"])
def test_table_layout_variants(emul, budget, monkeypatch):
    """VPT_SEED_BUDGET shrinks the shared-memory seed budget so that small models exercise the fat-bucket layout
    and the dense 16-bit-seed layout of large dictionaries."""
    from vpt_testlib import synth
    monkeypatch.setenv("VPT_SEED_BUDGET", budget)
    mb = synth.gen_model_bccwj_shaped(n_patterns=3000, sample_sentences=5000, dict_words=800)
    o = OraclePredictor(mb)
    text, offs, _ = synth.gen_text(60, ragged=True)
    for i in range(60):
        s = bytes(text[int(offs[i]):int(offs[i + 1])]).decode()
        sc, _, _, _ = run(emul, mb, s)
        assert sc == o.predict(s)[0].tolist()


def test_builder_survives_mutated_models(tmp_path):
    """ASan + UBSan build of the model reader + host predictor builder against mutated model files (random bytes, bit
    flips, small values in length / window positions): built or rejected with an error, never a crash.  Tag predictors
    also get their flat tag tables built, and the tag kernels' per-token code (host build) runs over them with random
    pattern-id states."""
    import struct
    exe = str(tmp_path / "build_fuzz")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                           "-I" + CSRC, "-I" + os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include"),
                           os.path.join(HERE, "native", "build_fuzz.cpp")] +
                          [os.path.join(CSRC, f) for f in ("model.cpp", "builder.cpp", "predictor_build.cpp", "tags_build.cpp")] +
                          ["-o", exe])
    samples = [open(os.path.join(GOLDEN, fn), "rb").read() for fn in ("model.bin", "tantivy_model.bin")]
    samples += [encode_model(m) for m in (kat.PREDICTOR_TEST_MODEL, kat.CHAR_ADD_SCORES_WITH_TAGS["model"],
                                          kat.CHAR_ADD_SCORES_3["model"], kat.TYPE_ADD_SCORES["model"],
                                          kat.TYPE_ADD_SCORES_WITH_TAGS["model"])]
    with open(tmp_path / "samples.bin", "wb") as f:
        for b in samples:
            f.write(struct.pack("<I", len(b)) + b)
    out = subprocess.run([exe, str(tmp_path / "samples.bin"), "3000"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "build fuzz done" in out.stdout


# ---- tag tables (tags_build.cpp: token table, key lists, chain tables) interpreted with the kernels' steps ---------------

def _fullwidth(text):
    import vaporetto_b200 as vb
    return "".join(chr(vb.lib().vpt_kytea_fullwidth(ord(c))) for c in text)


def run_tags(L, model_bytes, text, o, norm=False):
    """Tag prediction from the product's flat tag tables on the CPU, by the kernels' own per-token code compiled for the
    host (csrc/tags_token.hpp through tests/native/host_emul.cpp: emul_predict_tags), on the oracle's boundaries and the
    emulated pattern-id states; returns (tag_token, tag_cand[n, n_tags], unserved).  norm: the CLI's pre-filter -- the
    scorer sees the full-width image of the text, the tokens are looked up by the image of their ORIGINAL bytes."""
    seen = _fullwidth(text) if norm else text
    _, bd = o.predict(seen)
    _, cs, ts, _ = run(L, model_bytes, seen, tags=True)
    b = text.encode()
    n = len(cs)
    nt = max(o.n_tags, 1)
    bd = np.ascontiguousarray(np.asarray(bd, np.uint8))
    cs = np.asarray(cs, np.uint32)
    ts = np.asarray(ts, np.uint32)
    tok = np.zeros(n, np.int32)
    cand = np.zeros(n * nt, np.int32)
    uns = np.zeros(1, np.int32)
    rc = L.emul_predict_tags(model_bytes, len(model_bytes), b, len(b), bd.ctypes.data, cs.ctypes.data, ts.ctypes.data,
                             int(norm), tok.ctypes.data, cand.ctypes.data, uns.ctypes.data)
    assert rc == o.n_tags, L.emul_last_error()
    return tok, cand.reshape(n, nt)[:, :o.n_tags], int(uns[0])


def _check_tags(L, mb, texts, norm=False):
    o = OraclePredictor(mb, predict_tags=True)
    for text in texts:
        tok, cand, uns = run_tags(L, mb, text, o, norm=norm)
        ott, oti = o.predict_tags(_fullwidth(text) if norm else text)
        assert uns == 0
        # (token ids may be numbered differently only if tokens repeat in the model; the candidates must agree)
        assert (tok >= 0).tolist() == (ott >= 0).tolist(), text
        assert cand.tolist() == oti.tolist(), text


def test_tag_tables_on_reference_vectors(emul):
    # predictor.rs:863-903 (tags of "この人は地球人だ") and the bundled model's doctest sentences
    _check_tags(emul, encode_model(kat.PREDICTOR_TEST_MODEL), ["この人は地球人だ", "地球人", "この人"])
    with open(os.path.join(GOLDEN, "model.bin"), "rb") as f:
        _check_tags(emul, f.read(), ["まぁ社長は火星猫だ", "まぁ良いだろう", "火星", "社長は社長だ" * 5])


@pytest.mark.parametrize("cw,tw,maxdict,tags", [(3, 3, 5, 3), (1, 5, 3, 2), (2, 2, 2, 4), (3, 3, 9, 6), (5, 4, 4, 3)])
def test_tag_tables_random_models(emul, cw, tw, maxdict, tags):
    """Random tag models (suffix chains through n-grams and dictionary words, several rel positions, windows wider than
    three): the key-list scan with the truncation rule gives the sums of the reference's build-time merge."""
    from test_gpu_parity import _random_model
    rng = np.random.default_rng(177 + 1000 * cw + 100 * tw + maxdict + tags)
    for _ in range(4):
        model, alpha = _random_model(rng, cw, tw, maxdict=maxdict, tags=tags)
        texts = ["".join(rng.choice(list(alpha), size=rng.integers(1, 50))) for _ in range(60)]
        _check_tags(emul, encode_model(model), texts)


def test_tag_tables_synthetic_model(emul):
    from vpt_testlib import synth
    mb = synth.gen_model_bccwj_shaped(n_patterns=8000, sample_sentences=20000, tag_models=400)
    text, offs, _ = synth.gen_text(40, 40, seed=synth.TEXT_SEED + 13)
    _check_tags(emul, mb, [bytes(text[int(offs[i]):int(offs[i + 1])]).decode() for i in range(len(offs) - 1)])


def test_fused_probe_order_on_the_fuzz_case(emul):
    """emul_predict also walks the char table in k_fused's probe order (2-symbol node first, child mask, then the 3- or the
    1-symbol node) and fails when it selects another record than the longest-first order.  The model the GPU fuzzer found
    (a (parent node, symbol) record whose parent id equals a code point aliases an absent bigram unless the deep-key marker
    bits are compared: tests/golden/fuzz_cases) passes with the kernel's compare."""
    d = os.path.join(GOLDEN, "fuzz_cases")
    mb = np.load(os.path.join(d, "case1266_model.npy")).tobytes()
    text, offs = np.load(os.path.join(d, "case1266_text.npy")), np.load(os.path.join(d, "case1266_offs.npy"))
    o = OraclePredictor(mb)
    for i in range(len(offs) - 1):
        s = bytes(text[int(offs[i]):int(offs[i + 1])]).decode()
        sc, _, _, info = run(emul, mb, s)
        assert sc == o.predict(s)[0].tolist()


def test_tag_tables_with_the_fullwidth_prefilter(emul):
    """norm = 1 (the CLI default): tokens are looked up by the bytes of their KyteaFullwidthFilter image, computed on the fly
    from the original bytes (ASCII -> three-byte full-width forms: the byte length changes).  Tokens of the model are
    full-width strings; the text mixes ASCII, half-width punctuation and kana."""
    rng = np.random.default_rng(99)
    alpha = list("あいう人aB1x!?｡-ｱé𠀋Ω")  # (ｱ is not mapped: only the ten listed non-ASCII sources are; 2- and 4-byte characters)
    fw = _fullwidth
    tms = []
    for t in range(8):
        tok = fw("".join(rng.choice(alpha, size=rng.integers(1, 4))))
        cn = [(fw("".join(rng.choice(alpha, size=rng.integers(1, 3)))), [(int(rng.integers(0, 4)), rng.integers(-99, 99, size=2).tolist())])
              for _ in range(4)]
        tms.append(dict(token=tok, tags=[["x", "y"]], char_ngrams=cn, type_ngrams=[], bias=[1, 2]))
    cng = {fw("".join(rng.choice(alpha, size=rng.integers(1, 4)))): rng.integers(-500, 500, size=4).tolist() for _ in range(30)}
    model = dict(char_ngrams=list(cng.items()), type_ngrams=[(bytes([2]), [5, -5, 7, 1, 0, 2])], dict=[], bias=-3,
                 char_window=3, type_window=3, tag_models=tms)
    texts = ["".join(rng.choice(alpha, size=rng.integers(1, 30))) for _ in range(80)]
    _check_tags(emul, encode_model(model), texts, norm=True)
    _check_tags(emul, encode_model(model), [fw(t) for t in texts], norm=False)
