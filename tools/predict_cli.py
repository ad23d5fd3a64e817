#!/usr/bin/env python
"""The reference's `predict` command (predict/src/main.rs) on top of vpt_tokenize_lines: stdin lines -> space-separated
tokens on stdout, everything between the two (line splitting, full-width pre-filter, scoring, --wsconst post-filters,
output text) on the GPU.

    python tools/predict_cli.py --model model.bin[.zst] [--no-norm] [--wsconst D] [--wsconst R] ... < in.txt > out.txt

Options not on the device path (--scores, --tag-scores) are rejected; use the Sentence API
(vaporetto_b200.Sentence / include/vaporetto_b200.hpp) for tags."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def read_model(path: str) -> bytes:
    """The CLI reads a zstd-compressed model (main.rs:110-111); the library decodes it (vpt_model_read_zstd)."""
    with open(path, "rb") as f:
        return f.read()


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="A program to perform word segmentation (vaporetto_b200).")
    ap.add_argument("--model", required=True, help="The model file to use when analyzing text")
    ap.add_argument("--wsconst", action="append", default=[], choices=list("DRHTKOG"),
                    help="Do not segment some character types: D Digit, R Roman, H Hiragana, T Katakana, K Kanji, O Other, G Grapheme cluster")
    ap.add_argument("--predict-tags", action="store_true", help="Predicts POS tags")
    ap.add_argument("--no-norm", action="store_true", help="Do not normalize input strings before prediction")
    ap.add_argument("--device", type=int, default=0, help="CUDA device ordinal")
    args = ap.parse_args(argv)

    import vaporetto_b200 as vb
    print("Loading model file...", file=sys.stderr)
    predictor = vb.Predictor(vb.Model.read_zstd(read_model(args.model)), predict_tags=args.predict_tags, device=args.device)
    print("Start tokenization", file=sys.stderr)
    data = sys.stdin.buffer.read()
    t0 = time.perf_counter()
    out, _ = predictor.tokenize_lines(data, no_norm=args.no_norm, wsconst="".join(args.wsconst), predict_tags=args.predict_tags)
    dt = time.perf_counter() - t0
    sys.stdout.buffer.write(out.tobytes())
    print(f"Elapsed: {dt} [sec]", file=sys.stderr)
    return 0


if __name__ == "__main__":
    sys.exit(main())
