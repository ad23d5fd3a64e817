"""Builds and runs the C++ mirror (include/vaporetto_b200.hpp) test program: host parts on CPU, the
reference's doctests on the GPU."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EXE = os.path.join(HERE, "native", "cpp_mirror_test")
SRC = os.path.join(HERE, "native", "cpp_mirror_test.cpp")
MODEL = os.path.join(HERE, "golden", "model.bin")


def build():
    deps = [SRC, os.path.join(ROOT, "include", "vaporetto_b200.hpp"), os.path.join(ROOT, "include", "vaporetto_b200.h")]
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-o", EXE, SRC,
                               "-L" + os.path.join(ROOT, "vaporetto_b200"), "-lvaporetto_b200",
                               "-Wl,-rpath," + os.path.join(ROOT, "vaporetto_b200")])
    return EXE


def test_cpp_mirror_host():
    out = subprocess.run([build(), MODEL], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "ok" in out.stdout


@pytest.mark.gpu
def test_cpp_mirror_gpu():
    out = subprocess.run([build(), MODEL, "gpu"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "cpp mirror (gpu) ok" in out.stdout
