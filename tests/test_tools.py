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
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import predict_cli
    import pyarrow as pa
    raw = open(MODEL, "rb").read()
    z = tmp_path / "model.bin.zst"
    with pa.output_stream(str(z), compression="zstd") as s:  # the CLI's model files are zstd streams (main.rs:110)
        s.write(raw)
    assert predict_cli.read_model(str(z)) == raw
    assert predict_cli.read_model(MODEL) == raw


def test_cli_rejects_unsupported_options():
    out = subprocess.run([sys.executable, CLI, "--model", MODEL, "--wsconst", "G"], input=b"", capture_output=True)
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
