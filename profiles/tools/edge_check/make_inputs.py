#!/usr/bin/env python
"""Inputs of profiles/tools/edge_check/edge_check.cpp (a GPU check that needs no Python on the box: 1.6 s including CUDA
start-up): for two model shapes -- (3, 3): k_fused, common shape; (6, 3): general rows through k_tile_fast -- a random model
whose type n-grams all carry weights, 302 sentences over both sides of every CharacterType range edge, and the scores the CPU
oracle gives them.  Writes build/edge/c<cw><tw>.{model,text,offs,scores}; build and run:
  g++ -O1 -std=c++17 -o build/edge/edge_check profiles/tools/edge_check/edge_check.cpp -Lvaporetto_b200 -lvaporetto_b200 \\
      -Wl,-rpath,'$ORIGIN/../../vaporetto_b200'
  gpurun -- 'build/edge/edge_check build/edge/c33 build/edge/c63'"""
import importlib.util
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from vpt_testlib.bincode_model import encode_model  # noqa: E402
from vpt_testlib.oracle import OraclePredictor  # noqa: E402

spec = importlib.util.spec_from_file_location("tgp", os.path.join(ROOT, "tests", "test_gpu_parity.py"))
tgp = importlib.util.module_from_spec(spec)
spec.loader.exec_module(tgp)
src = open(os.path.join(ROOT, "tests", "test_gpu_parity.py")).read()
edges = eval(re.search(r"edges = (\[.*?\])\n    rng", src, re.S).group(1))
os.makedirs(os.path.join(ROOT, "build", "edge"), exist_ok=True)
for cw, tw in [(3, 3), (6, 3)]:
    rng = np.random.default_rng(4242 + 10 * cw + tw)
    model, alpha = tgp._random_model(rng, cw, tw, maxdict=3)
    tng = {}
    for n in (1, 2, 3):
        for k in range(6 ** n):
            g = bytes(1 + (k // 6 ** j) % 6 for j in range(n))
            tng[g] = rng.integers(-32767, 32768, size=max(2 * tw - n + 1, 0)).tolist()
    model["type_ngrams"] = list(tng.items())
    mb = encode_model(model)
    o = OraclePredictor(mb)
    pool = [chr(c) for c in edges] + list(alpha)
    sents = ["".join(rng.choice(pool, size=rng.integers(2, 70))) for _ in range(300)]
    sents += ["".join(chr(c) for c in edges), "".join(chr(c) for c in range(0x4DB0, 0x4E10))]
    enc = [s.encode() for s in sents]
    offs = np.zeros(len(enc) + 1, np.uint64)
    np.cumsum([len(e) for e in enc], out=offs[1:])
    text = np.frombuffer(b"".join(enc), np.uint8)
    sc, bd, boff, st = o.predict_batch(text, offs, nthreads=2)
    assert (st == 0).all()
    base = os.path.join(ROOT, "build", "edge", f"c{cw}{tw}")
    open(base + ".model", "wb").write(mb)
    text.tofile(base + ".text")
    offs.tofile(base + ".offs")
    sc.astype(np.int32).tofile(base + ".scores")
    print(base, len(sents), "sentences,", len(sc), "boundaries")
