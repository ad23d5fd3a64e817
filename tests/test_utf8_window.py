"""The 4-byte window decoders of k_fused's stream stage (vaporetto_b200/csrc/utf8_window.hpp) on the host: every window
against a byte-by-byte restatement of str::from_utf8's rules (reference sentence.rs:160-196), and the fast path for
ASCII / three-byte leads against the general decoder; char_type and the tile kernels' BMP type table (textnorm.hpp) against
the reference's ranges for every code point (tests/native/utf8_window_test.cpp; 3.6e8 windows, a few seconds)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "native", "utf8_window_test.cpp")
EXE = os.path.join(HERE, "native", "utf8_window_test")
HDRS = [os.path.join(ROOT, "vaporetto_b200", "csrc", f) for f in ("utf8_window.hpp", "textnorm.hpp")]


def test_window_decoders_exhaustive():
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(f) for f in [SRC] + HDRS):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", EXE, SRC])
    out = subprocess.run([EXE], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("utf8 window ok")
