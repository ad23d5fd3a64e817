"""KyTea binary model -> vaporetto model (reference kytea_model.rs; SURVEY §8(f) rank 4, "KyTea converter"):
the library's single-pass converter (csrc/kytea_model.cpp) and `Model::to_vec` against the oracle's struct-by-struct
restatement, the reference's doctest on its own fixture, and the host tables built from the converted model."""
import numpy as np
import os
import pytest

import vaporetto_b200 as vb
from vpt_testlib import kytea_writer as kw
from vpt_testlib.oracle import OraclePredictor, kytea_to_model_bytes
from test_host_tables import emul, run  # noqa: F401  (fixture + helper)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def read(fn):
    with open(os.path.join(GOLDEN, fn), "rb") as f:
        return f.read()


def test_to_vec_round_trips_reference_written_files():
    # files written by the reference's Model::write: reading and writing them back must give the same bytes
    for fn in ("model.bin", "tantivy_model.bin"):
        raw = read(fn)
        assert vb.Model.read(raw).to_vec() == raw


def test_kytea_fixture_doctest(emul):  # noqa: F811
    """kytea_model.rs:401-422: resources/kytea-model.bin -> Model -> Predictor(model, false) tokenises
    "まぁ社長は火星猫だ" as "まぁ 社長 は 火星 猫 だ"."""
    k = read("kytea-model.bin")
    ours = vb.Model.read_kytea(k).to_vec()
    assert ours == kytea_to_model_bytes(k)
    o = OraclePredictor(ours)
    assert o.tokenize("まぁ社長は火星猫だ") == "まぁ 社長 は 火星 猫 だ"
    # the device tables built from the converted model, interpreted on the CPU with the kernels' probe sequence
    for text in ("まぁ社長は火星猫だ", "まぁ良いだろう", "火星猫の生態を調査した2021年", "a"):
        sc, _, _, _ = run(emul, ours, text)
        assert sc == o.predict(text)[0].tolist()
    # a host-only predictor can be built from it (no GPU here; scoring needs the device)
    p = vb.Predictor(vb.Model.read_kytea(k), device=-1)
    assert p.info["char_window"] > 0


def test_random_kytea_models_match_oracle(emul):  # noqa: F811
    rng = np.random.default_rng(20240924)
    accepted = 0
    for it in range(150):
        k = kw.random_model(rng)
        want = kytea_to_model_bytes(k)
        got = vb.Model.read_kytea(k).to_vec()
        assert got == want, it
        if it % 5 == 0:
            try:
                o = OraclePredictor(want)
            except Exception:
                continue  # Predictor::new rejects some random models (e.g. a weight vector longer than its window)
            accepted += 1
            for text in ("あいう火星猫aB1。é", "火星猫" * 5 + "\U00020000a1"):
                sc, _, _, _ = run(emul, want, text)
                assert sc == o.predict(text)[0].tolist(), it
    assert accepted >= 5


def test_kytea_errors():
    w = kw.KyteaWriter("aDRHTKO\x05")
    lk = dict(chars={"a": [1, 2, 3]}, types={"D": [1, 2, 3]}, dict_vec=[0] * 3, biases=[7])
    ok = w.model(char_w=1, type_w=1, dict_n=1, wordseg=dict(lookup=lk))
    assert vb.Model.read_kytea(ok).to_vec() == kytea_to_model_bytes(ok)
    cases = {
        "no word segmentation model.": w.model(char_w=1, type_w=1, dict_n=1, wordseg=None),
        "no lookup data.": w.model(char_w=1, type_w=1, dict_n=1, wordseg=dict(lookup=None)),
        "no character dictionary.": w.model(char_w=1, type_w=1, dict_n=1, wordseg=dict(lookup=dict(lk, chars={}))),
        "no type dictionary.": w.model(char_w=1, type_w=1, dict_n=1, wordseg=dict(lookup=dict(lk, types={}))),
        "unsupported character type: 5": w.model(char_w=1, type_w=1, dict_n=1,
                                                 wordseg=dict(lookup=dict(lk, types={"\x05": [1, 2, 3]}))),
    }
    for msg, data in cases.items():
        with pytest.raises(vb.VaporettoError) as e:
            vb.Model.read_kytea(data)
        assert msg in str(e.value), (msg, str(e.value))
        with pytest.raises(Exception) as e2:
            kytea_to_model_bytes(data)
        assert msg in str(e2.value)
    for cut in (0, 5, 30, len(ok) // 2, len(ok) - 1):  # truncated files: an I/O error, never a crash
        with pytest.raises(vb.VaporettoError):
            vb.Model.read_kytea(ok[:cut])
        with pytest.raises(Exception):
            kytea_to_model_bytes(ok[:cut])


def test_dictionary_access_and_replacement():
    """Model::dictionary / replace_dictionary (model.rs:155-163) + WordWeightRecord::new's check (dict_model.rs:39-50):
    with Model::to_vec this is the reference's manipulate_model tool; bytes compared with the test-side bincode encoder."""
    from vpt_testlib.bincode_model import encode_model
    spec = dict(char_ngrams=[("火星", [1, 2, 3])], type_ngrams=[(bytes([5, 5]), [4, 5, 6])],
                dict=[("火星猫", [1, -2, 3, 4], "comment"), ("é", [5, 6], "")], bias=7, char_window=2, type_window=2,
                tag_models=[])
    m = vb.Model.read(encode_model(spec))
    assert m.dictionary() == [("火星猫", [1, -2, 3, 4], "comment"), ("é", [5, 6], "")]
    new = [("社長", [10, 20, -30], "x"), ("\U00020000", [1, 2], "")]
    m.replace_dictionary(new)
    assert m.dictionary() == new
    assert m.to_vec() == encode_model(dict(spec, dict=new))
    with pytest.raises(vb.VaporettoError) as e:
        m.replace_dictionary([("社長", [1, 2], "")])   # needs 3 weights
    assert "does not match the length of the `word`" in str(e.value)
    assert m.dictionary() == new                        # unchanged after the failed call
    m.replace_dictionary([])
    assert m.dictionary() == [] and m.to_vec() == encode_model(dict(spec, dict=[]))
    # the converted KyTea fixture exposes its dictionary: [left, inside.., right] per word
    km = vb.Model.read_kytea(read("kytea-model.bin"))
    for word, weights, _ in km.dictionary():
        assert len(weights) == len(word) + 1


def test_parser_survives_mutated_files(tmp_path):
    """ASan + UBSan build of the parser against 20 000 mutated KyTea files (bit flips, overwritten counts, truncation):
    every file is either converted (and its written form reads back identically) or rejected with an error."""
    import struct
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    csrc = os.path.join(os.path.dirname(here), "vaporetto_b200", "csrc")
    exe = str(tmp_path / "kytea_fuzz")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                           "-I" + csrc, os.path.join(here, "native", "kytea_fuzz.cpp"), os.path.join(csrc, "kytea_model.cpp"),
                           os.path.join(csrc, "model.cpp"), "-o", exe])
    rng = np.random.default_rng(11)
    samples = [read("kytea-model.bin")] + [kw.random_model(rng) for _ in range(30)]
    with open(tmp_path / "samples.bin", "wb") as f:
        for b in samples:
            f.write(struct.pack("<I", len(b)) + b)
    out = subprocess.run([exe, str(tmp_path / "samples.bin"), "20000"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "fuzz done" in out.stdout
    # the native model reader (Model::read, bincode) under the same treatment
    from golden import reference_kat as kat
    from vpt_testlib.bincode_model import encode_model
    native = [read("model.bin"), read("tantivy_model.bin"), encode_model(kat.PREDICTOR_TEST_MODEL),
              encode_model(kat.CHAR_ADD_SCORES_WITH_TAGS["model"])]
    with open(tmp_path / "native.bin", "wb") as f:
        for b in native:
            f.write(struct.pack("<I", len(b)) + b)
    out = subprocess.run([exe, str(tmp_path / "native.bin"), "20000", "native"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "fuzz done" in out.stdout
