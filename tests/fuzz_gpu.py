#!/usr/bin/env python
"""Randomised GPU-vs-oracle parity fuzzer (run manually on the GPU box: python tests/fuzz_gpu.py [seconds]).
Random models (windows 0-5, n-grams 1-4 chars, dictionary words up to 20 chars, optional tag models, duplicate
entries, short/over-long weight vectors) x random batches (lengths 1-600, mixed scripts, 4-byte characters).
Every model also goes through vpt_tokenize_lines (both with and without the full-width pre-filter) on the same
sentences joined by random line terminators, with empty / malformed lines mixed in and a random chunk size."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vaporetto_b200 as vb  # noqa: E402
from vpt_testlib.bincode_model import encode_model  # noqa: E402
from vpt_testlib.oracle import OraclePredictor  # noqa: E402

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
FAIL_MODEL = os.path.join(ROOT, "gpurun_out", "fuzz_fail_model.npy")   # (gpurun_out/ travels back from the GPU box)
FAIL_LINES = os.path.join(ROOT, "gpurun_out", "fuzz_fail_lines.bin")
ALPHA = list("あいうえおかきアイウエ人火星地球猫社長漢字aBc1 9。、🤌𠀋é")
# both sides of the edges of CharacterType::get_type's ranges that lie inside a 256-code-point page, and of the encoding
# lengths (the tile kernels type the BMP from a page table: a page with two types needs a sub-table; U+4DBF | U+4DC0 had none)
ALPHA += [chr(c) for c in (0x3096, 0x3097, 0x30FA, 0x30FB, 0x4DBF, 0x4DC0, 0x9FFF, 0xA000, 0xFAFF, 0xFB00, 0xFF19, 0xFF1A, 0xFF9F,
                           0xFFA0, 0x7F, 0x80, 0x7FF, 0x800, 0xFFFF, 0x10000, 0x2A6DF, 0x2A6E0, 0x10FFFF)]
LINE_EXTRA = list("/\\.-ａ１。－―｢")  # escapes and sources / targets of the full-width filter


def rand_model(rng):
    cw, tw = int(rng.integers(0, 6)), int(rng.integers(0, 6))
    word = lambda lo, hi: "".join(rng.choice(ALPHA, size=rng.integers(lo, hi + 1)))
    wv = lambda n: rng.integers(-40000, 40000, size=max(int(n), 0)).tolist()
    cng = [(word(1, 4), wv(2 * cw - rng.integers(0, 4) + rng.integers(0, 3))) for _ in range(rng.integers(0, 60))]
    dic = [(w, wv(len(w) + rng.integers(-1, 3)), "") for w in (word(1, 20 if rng.random() < 0.2 else 6) for _ in range(rng.integers(0, 60)))]
    tng, seen = [], set()
    for _ in range(rng.integers(0, 40)):
        g = bytes(rng.integers(1, 7, size=rng.integers(1, 5)).tolist())
        if g in seen and tw <= 3:
            continue  # the cache variant rejects duplicate type n-grams
        seen.add(g)
        tng.append((g, wv(2 * tw - len(g) + 1 + rng.integers(-1, 2))))
    tms, toks = [], set()
    for t in range(rng.integers(0, 4) if rng.random() < 0.5 else 0):
        tok = word(1, 3)
        if tok in toks:
            continue
        toks.add(tok)
        nc = [int(rng.integers(1, 5)) for _ in range(rng.integers(1, 3))]
        sl = sum(c for c in nc if c >= 2)
        tms.append(dict(token=tok, tags=[["t%d_%d" % (k, j) for j in range(c)] for k, c in enumerate(nc)],
                        char_ngrams=[(word(1, 4), [(int(rng.integers(0, cw + 1)), wv(sl + rng.integers(0, 2)))]) for _ in range(rng.integers(0, 6))],
                        type_ngrams=[(bytes(rng.integers(1, 7, size=rng.integers(1, 5)).tolist()), [(int(rng.integers(0, tw + 1)), wv(sl))]) for _ in range(rng.integers(0, 4))],
                        bias=wv(sl + rng.integers(0, 3))))
    return dict(char_ngrams=cng, type_ngrams=tng, dict=dic, bias=int(rng.integers(-50000, 50000)), char_window=cw,
                type_window=tw, tag_models=tms)


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    rng = np.random.default_rng(int(os.environ.get("FUZZ_SEED", "12345")))
    t0, it, paths = time.time(), 0, {}
    while time.time() - t0 < budget:
        it += 1
        model = rand_model(rng)
        tags = bool(model["tag_models"]) and rng.random() < 0.8
        mb = encode_model(model)
        try:
            o = OraclePredictor(mb, predict_tags=tags)
        except Exception as e:  # models the reference rejects must be rejected here too
            try:
                vb.Predictor(vb.Model.read(mb), predict_tags=tags)
                raise SystemExit(f"iteration {it}: oracle rejected the model ({e}) but the product accepted it")
            except vb.VaporettoError:
                continue
        p = vb.Predictor(vb.Model.read(mb), predict_tags=tags)
        key = (p.info["fast_path"], p.info["char_scorer"], p.info["type_scorer"])
        paths[key] = paths.get(key, 0) + 1
        lens = rng.integers(1, 120, size=rng.integers(1, 400))
        if rng.random() < 0.3:
            lens[rng.integers(0, len(lens))] = rng.integers(600, 4000)
        sents = ["".join(rng.choice(ALPHA, size=n)) for n in lens]
        blob = "".join(sents).encode()
        offs = np.zeros(len(sents) + 1, np.uint64)
        np.cumsum([len(s.encode()) for s in sents], out=offs[1:])
        text = np.frombuffer(blob, np.uint8)
        r = p.predict_batch(text, offs, want_states=tags)
        sc, bd, boff, st = o.predict_batch(text, offs, nthreads=4)
        if not (np.array_equal(r.scores, sc) and np.array_equal(r.boundaries, bd) and np.array_equal(r.bound_offsets, boff)):
            np.save(FAIL_MODEL, np.frombuffer(mb, np.uint8))
            np.save(os.path.join(ROOT, "gpurun_out", "fuzz_fail_text.npy"), text)
            np.save(os.path.join(ROOT, "gpurun_out", "fuzz_fail_offs.npy"), offs)
            r_again = p.predict_batch(text, offs, want_states=tags)
            print("second run of the same batch equal to the first:", np.array_equal(r_again.scores, r.scores),
                  "equal to the oracle:", np.array_equal(r_again.scores, sc), file=sys.stderr)
            raise SystemExit(f"iteration {it}: MISMATCH (model saved to gpurun_out/fuzz_fail_model.npy), path {key}")
        # the CLI loop on the device (untagged output): the same sentences as lines
        parts = []
        for sline in sents[: 120]:
            r2 = rng.random()
            if r2 < 0.05:
                body = b""
            elif r2 < 0.08:
                body = bytes(rng.integers(0, 256, rng.integers(1, 9)).astype(np.uint8)).replace(b"\n", b"")
            elif r2 < 0.4:
                body = (sline[: 1 + len(sline) // 2] + "".join(rng.choice(LINE_EXTRA, size=rng.integers(1, 6)))).encode()
            else:
                body = sline.encode()
            parts.append(body + (b"\r\n" if rng.random() < 0.15 else b"\n"))
        data = b"".join(parts)
        if rng.random() < 0.5 and data.endswith(b"\n") and not data.endswith(b"\r\n"):
            data = data[:-1]
        os.environ["VPT_CHUNK_BYTES"] = str(int(rng.choice([64, 777, 1 << 14, 16 << 20])))
        for no_norm in (True, False):
            ws = "".join(rng.choice(list("DRHTKOG"), size=rng.integers(0, 3)))  # --wsconst options
            with_tags = tags and rng.random() < 0.7                              # --predict-tags
            try:
                got, nl = p.tokenize_lines(data, no_norm=no_norm, wsconst=ws, predict_tags=with_tags)
            except vb.VaporettoError as e:
                if with_tags and e.code == 17:   # tag model beyond the device limits: the host path serves it
                    continue
                raise
            want, wl = o.tokenize_lines(data, no_norm=no_norm, wsconst=ws, predict_tags=with_tags)
            if nl != wl or got.tobytes() != want:
                np.save(FAIL_MODEL, np.frombuffer(mb, np.uint8))
                open(FAIL_LINES, "wb").write(data)
                raise SystemExit(f"iteration {it}: tokenize_lines MISMATCH (no_norm={no_norm}, wsconst={ws!r}, tags={with_tags}), path {key}")
        # compact results: the bit stream against the byte boundaries; the token records against the oracle's tags
        try:
            cr = p.predict_batch_compact(text, offs, tags=tags)
        except vb.VaporettoError as e:
            if not (tags and e.code == 17):
                raise
            cr = p.predict_batch_compact(text, offs, tags=False)
        if not (np.array_equal(cr.boundaries(), bd) and np.array_equal(cr.status.astype(np.int32) == 0, st == 0)):
            raise SystemExit(f"iteration {it}: compact boundaries MISMATCH, path {key}")
        if tags and cr.token_ids is not None and cr.n_unserved == 0:
            for i in rng.integers(0, len(sents), size=3):
                if st[i] != 0:
                    continue
                try:
                    ott, oti = o.predict_tags(sents[i])
                except Exception:
                    continue   # a tag model the reference rejects at prediction time
                lo, hi = int(cr.token_offsets[i]), int(cr.token_offsets[i + 1])
                ends = np.nonzero(np.concatenate([cr.boundaries(int(i)) == 1, [True]]))[0]
                if hi - lo != len(ends):
                    raise SystemExit(f"iteration {it}: token count mismatch, path {key}")
                got_known = cr.token_ids[lo:hi] >= 0
                got_c = np.where(cr.token_cands[lo:hi] == 255, -1, cr.token_cands[lo:hi].astype(np.int64))
                if not (np.array_equal(got_known, ott[ends] >= 0) and np.array_equal(got_c[got_known], oti[ends][got_known])):
                    raise SystemExit(f"iteration {it}: token records MISMATCH (sentence {i}), path {key}")
        if tags:
            for i in rng.integers(0, len(sents), size=3):
                _, _, ocs, ots = o.predict(sents[i], states=True)
                c0 = int(r.char_offsets[i])
                if p.info["char_scorer"] == 2 and not np.array_equal(r.char_states[c0:c0 + len(ocs)], ocs):
                    raise SystemExit(f"iteration {it}: char state mismatch, path {key}")
                if p.info["type_scorer"] == 3 and not np.array_equal(r.type_states[c0:c0 + len(ots)], ots):
                    raise SystemExit(f"iteration {it}: type state mismatch, path {key}")
    print(f"fuzz ok: {it} models in {time.time() - t0:.0f}s; (fast_path, char_scorer, type_scorer) counts: {paths}")


if __name__ == "__main__":
    main()
