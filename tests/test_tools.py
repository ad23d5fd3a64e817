"""tools/predict_cli.py: the reference's `predict` command line on top of vpt_tokenize_lines."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CLI = os.path.join(ROOT, "tools", "predict_cli.py")
MODEL = os.path.join(HERE, "golden", "model.bin")


def test_read_model_raw_and_zstd(tmp_path):
    """vpt_model_read_zstd (the CLI's model loading, predict/src/main.rs:110-111): streamed zstd frames (no stored content
    size), concatenated frames and raw images give the model the raw image gives; damaged streams fail with IOError."""
    import pyarrow as pa
    import vaporetto_b200 as vb
    raw = open(MODEL, "rb").read()
    z = tmp_path / "model.bin.zst"
    with pa.output_stream(str(z), compression="zstd") as s:  # the CLI's model files are zstd streams (main.rs:110)
        s.write(raw)
    zb = open(z, "rb").read()
    assert zb[:4] == b"\x28\xb5\x2f\xfd" and zb != raw
    want = vb.Model.read(raw).to_vec()
    assert vb.Model.read_zstd(zb).to_vec() == want
    assert vb.Model.read_zstd(raw).to_vec() == want
    # two frames back to back decode as one stream (zstd::Decoder reads them all)
    half = len(raw) // 2
    parts = []
    for chunk in (raw[:half], raw[half:]):
        sink = pa.BufferOutputStream()
        with pa.CompressedOutputStream(sink, "zstd") as s:
            s.write(chunk)
        parts.append(sink.getvalue().to_pybytes())
    assert vb.Model.read_zstd(parts[0] + parts[1]).to_vec() == want
    # a large image (beyond the first output buffer)
    big = pa.compress(b"\0" * (9 << 20), codec="zstd", asbytes=True)
    with pytest.raises(vb.VaporettoError):
        vb.Model.read_zstd(big)  # decodes, then fails as a model (not as an I/O error)
    for bad in (zb[: len(zb) // 2], zb[:8] + b"\xff" * 16 + zb[24:]):
        with pytest.raises(vb.VaporettoError) as e:
            vb.Model.read_zstd(bad)
        assert e.value.code in (1, 4, 5)


def test_cli_rejects_unsupported_options():
    out = subprocess.run([sys.executable, CLI, "--model", MODEL, "--wsconst", "X"], input=b"", capture_output=True)
    assert out.returncode != 0 and b"invalid choice" in out.stderr


@pytest.mark.gpu
def test_cli_end_to_end():
    from vpt_testlib.oracle import OraclePredictor
    o = OraclePredictor(open(MODEL, "rb").read())
    text = "まぁ社長は火星猫だ\r\n\nまぁ良いだろう\nVaporetto 1.5\n".encode()
    out = subprocess.run([sys.executable, CLI, "--model", MODEL], input=text, capture_output=True)
    assert out.returncode == 0, out.stderr.decode()
    assert out.stdout == o.tokenize_lines(text)[0]
    assert out.stdout.decode().split("\n")[:3] == ["まぁ 社長 は 火星 猫 だ", "", "まぁ 良い だろう"]
    # --wsconst R --wsconst D --no-norm
    out2 = subprocess.run([sys.executable, CLI, "--model", MODEL, "--wsconst", "R", "--wsconst", "D", "--no-norm"],
                          input=text, capture_output=True)
    assert out2.returncode == 0, out2.stderr.decode()
    assert out2.stdout == o.tokenize_lines(text, no_norm=True, wsconst="RD")[0]
