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
