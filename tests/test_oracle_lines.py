"""The oracle's restatement of the reference CLI loop (`predict --no-norm`, predict/src/main.rs:126-150) against
the per-sentence oracle (itself pinned to the reference's vectors) and Rust's `BufRead::lines` rules as Python's
own line handling states them."""
import os

from vpt_testlib.oracle import OraclePredictor

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


def expected(o, data: bytes) -> bytes:
    out = []
    for line in rust_lines(data):
        try:
            s = line.decode("utf-8")
            ok = len(s) > 0 and "\x00" not in s
        except UnicodeDecodeError:
            ok = False
        out.append(o.tokenize(s).encode() if ok else b"")
    return b"".join(x + b"\n" for x in out)


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
    for data in cases:
        got, nl = o.tokenize_lines(data)
        assert nl == len(rust_lines(data))
        assert got == expected(o, data), data


def test_docs_tok_through_cli_loop():
    """The reference's documented tokenisation (tests/golden/docs.tok) through the whole-buffer loop."""
    with open(os.path.join(GOLDEN, "model.bin"), "rb") as f:
        o = OraclePredictor(f.read())
    got, nl = o.tokenize_lines("まぁ社長は火星猫だ\nまぁ社長は火星猫だ\n".encode())
    assert nl == 2
    assert got.decode() == "まぁ 社長 は 火星 猫 だ\n" * 2
