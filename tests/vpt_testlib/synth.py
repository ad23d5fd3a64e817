"""Synthetic Japanese-shaped text and models for the parity tests and bench.py.

Deterministic: counter-based SplitMix64 (seed 0x5EED0001 for text, 0x5EED0002 for models).
Character classes: hiragana .45 (U+3041..3096, Zipf s=1), kanji .33 (2136 code points from U+4E00, stride 9,
Zipf s=1), katakana .12 (U+30A1..30FA + U+30FC, Zipf s=1), ASCII digit .03, ASCII roman .03, other .04
(、。「」・！？ and space).  The class of a character follows a first-order Markov chain with 0.6 self-transition.
"""
from __future__ import annotations

import numpy as np

from .bincode_model import MODEL_MAGIC, varint, zigzag

TEXT_SEED = 0x5EED0001
MODEL_SEED = 0x5EED0002

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(seed: int, n: int, stream: int = 0) -> np.ndarray:
    """n outputs of SplitMix64 started at `seed` (counter-based, so vectorised), sub-stream `stream`."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed) + np.uint64(stream) * np.uint64(0xD1B54A32D192ED03) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _uniform(seed, n, stream):
    return (splitmix64(seed, n, stream) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def _zipf_cdf(k: int) -> np.ndarray:
    p = 1.0 / np.arange(1, k + 1)
    return np.cumsum(p / p.sum())


HIRAGANA = np.arange(0x3041, 0x3097, dtype=np.uint32)                        # 86
KANJI = (0x4E00 + 9 * np.arange(2136)).astype(np.uint32)                      # 2136
KATAKANA = np.concatenate([np.arange(0x30A1, 0x30FB), [0x30FC]]).astype(np.uint32)  # 91
DIGIT = np.arange(0x30, 0x3A, dtype=np.uint32)
ROMAN = np.concatenate([np.arange(0x61, 0x7B), np.arange(0x41, 0x5B)]).astype(np.uint32)
OTHER = np.array([ord(c) for c in "、。「」・！？ "], dtype=np.uint32)
CLASSES = [HIRAGANA, KANJI, KATAKANA, DIGIT, ROMAN, OTHER]
CLASS_P = np.array([0.45, 0.33, 0.12, 0.03, 0.03, 0.04])
CLASS_ZIPF = [True, True, True, False, False, False]
SELF_TRANSITION = 0.6


def gen_codepoints(n_sent: int, lengths, seed: int = TEXT_SEED) -> np.ndarray:
    """Returns a flat uint32 array of code points for sentences of the given lengths (int or array)."""
    lengths = np.full(n_sent, lengths, dtype=np.int64) if np.isscalar(lengths) else np.asarray(lengths, np.int64)
    total = int(lengths.sum())
    maxlen = int(lengths.max()) if n_sent else 0
    starts = np.zeros(n_sent + 1, np.int64)
    np.cumsum(lengths, out=starts[1:])
    cls = np.zeros(total, np.int8)
    class_cdf = np.cumsum(CLASS_P)
    prev = np.zeros(n_sent, np.int8)
    # position-major generation: random streams are indexed by (position, sentence)
    for pos in range(maxlen):
        alive = np.nonzero(lengths > pos)[0]
        u_stay = _uniform(seed, n_sent, 3 * pos + 0)[alive]
        u_cls = _uniform(seed, n_sent, 3 * pos + 1)[alive]
        fresh = np.searchsorted(class_cdf, u_cls, side="right").clip(0, 5).astype(np.int8)
        c = fresh if pos == 0 else np.where(u_stay < SELF_TRANSITION, prev[alive], fresh)
        prev[alive] = c
        cls[starts[alive] + pos] = c
    # character within class
    sent_of = np.repeat(np.arange(n_sent), lengths)
    pos_of = np.arange(total) - starts[sent_of]
    u_chr = np.empty(total)
    for pos in range(maxlen):
        sel = np.nonzero(pos_of == pos)[0]
        u_chr[sel] = _uniform(seed, n_sent, 3 * pos + 2)[sent_of[sel]]
    cps = np.zeros(total, np.uint32)
    for k, (tab, zipf) in enumerate(zip(CLASSES, CLASS_ZIPF)):
        sel = np.nonzero(cls == k)[0]
        if zipf:
            idx = np.searchsorted(_zipf_cdf(len(tab)), u_chr[sel], side="right").clip(0, len(tab) - 1)
        else:
            idx = (u_chr[sel] * len(tab)).astype(np.int64).clip(0, len(tab) - 1)
        cps[sel] = tab[idx]
    return cps


def encode_utf8(cps: np.ndarray, lengths: np.ndarray):
    """Flat code points + per-sentence char counts -> (uint8 text, uint64 byte offsets[n+1])."""
    cps = cps.astype(np.uint32)
    nb = np.where(cps < 0x80, 1, np.where(cps < 0x800, 2, np.where(cps < 0x10000, 3, 4))).astype(np.int64)
    ends = np.cumsum(nb)
    startb = ends - nb
    out = np.zeros(int(ends[-1]) if len(ends) else 0, np.uint8)
    m1 = nb == 1
    out[startb[m1]] = cps[m1]
    m2 = nb == 2
    out[startb[m2]] = 0xC0 | (cps[m2] >> 6)
    out[startb[m2] + 1] = 0x80 | (cps[m2] & 0x3F)
    m3 = nb == 3
    out[startb[m3]] = 0xE0 | (cps[m3] >> 12)
    out[startb[m3] + 1] = 0x80 | ((cps[m3] >> 6) & 0x3F)
    out[startb[m3] + 2] = 0x80 | (cps[m3] & 0x3F)
    m4 = nb == 4
    out[startb[m4]] = 0xF0 | (cps[m4] >> 18)
    out[startb[m4] + 1] = 0x80 | ((cps[m4] >> 12) & 0x3F)
    out[startb[m4] + 2] = 0x80 | ((cps[m4] >> 6) & 0x3F)
    out[startb[m4] + 3] = 0x80 | (cps[m4] & 0x3F)
    lengths = np.asarray(lengths, np.int64)
    cstart = np.zeros(len(lengths) + 1, np.int64)
    np.cumsum(lengths, out=cstart[1:])
    ends0 = np.concatenate([[0], ends])
    offsets = ends0[cstart].astype(np.uint64)
    return out, offsets


def gen_text(n_sent: int, length=40, seed: int = TEXT_SEED, ragged: bool = False):
    """Synthetic batch: (text uint8, offsets uint64[n+1], lengths).  ragged: clipped LogNormal(ln 35, 0.6) in [1,512]."""
    if ragged:
        u1 = _uniform(seed ^ 0xA5A5, n_sent, 1001)
        u2 = _uniform(seed ^ 0xA5A5, n_sent, 1002)
        z = np.sqrt(-2.0 * np.log(np.maximum(u1, 1e-300))) * np.cos(2 * np.pi * u2)
        lengths = np.clip(np.exp(np.log(35.0) + 0.6 * z), 1, 512).astype(np.int64)
    else:
        lengths = np.full(n_sent, length, np.int64)
    cps = gen_codepoints(n_sent, lengths, seed)
    text, offsets = encode_utf8(cps, lengths)
    return text, offsets, lengths


def _cp_to_utf8_bytes(cp: int) -> bytes:
    return chr(int(cp)).encode("utf-8")


def _ngram_counts(cps2d: np.ndarray, n: int):
    """Counts of the n-grams of every row of a [sentences, length] code-point matrix (within rows)."""
    L = cps2d.shape[1]
    key = np.zeros((cps2d.shape[0], L - n + 1), np.uint64)
    for k in range(n):
        key = (key << np.uint64(21)) | cps2d[:, k:L - n + 1 + k].astype(np.uint64)
    keys, counts = np.unique(key.ravel(), return_counts=True)
    return keys, counts


def _weights(seed, n_patterns, width, stream):
    """[n_patterns, width] weights: nonzero w.p. 0.5, uniform in [-32767, 32767] (trainer.rs:18,383)."""
    r = splitmix64(seed, n_patterns * width, stream)
    nz = (r & np.uint64(1)).astype(bool)
    mag = ((r >> np.uint64(8)) % np.uint64(65535)).astype(np.int64) - 32767
    return np.where(nz, mag, 0).astype(np.int64).reshape(n_patterns, width)


def _encode_ngram_list(ngram_bytes, weight_rows) -> bytes:
    """bincode Vec<NgramData{ngram: String|Vec<u8>, weights: Vec<i32>}> (fast path for big models)."""
    parts = [varint(len(ngram_bytes))]
    zz_cache = {}
    for g, ws in zip(ngram_bytes, weight_rows):
        parts.append(varint(len(g)))
        parts.append(g)
        parts.append(varint(len(ws)))
        for x in ws:
            x = int(x)
            e = zz_cache.get(x)
            if e is None:
                e = zigzag(x)
                zz_cache[x] = e
            parts.append(e)
    return b"".join(parts)


def gen_model_bccwj_shaped(n_patterns: int = 300_000, sample_sentences: int = 2_000_000, window: int = 3,
                           seed: int = MODEL_SEED, dict_words: int = 0, tag_models: int = 0) -> bytes:
    """bccwj-suw-shaped model (BASELINE config 2): the n_patterns most frequent char 1/2/3-grams of a synthetic
    sample, all 6+36+216 type 1-3-grams, W=3 both, no tags; `dict_words` > 0 adds a KyTea-shaped dictionary
    (config 4: lengths {1:2%,2:30%,3:25%,4:18%,5-7:17%,8-16:8%}, weights [L, I.., R] in 4 length buckets);
    `tag_models` > 0 adds unidic_pos-shaped tag models (config 3: tokens = 1-4-char strings, 1-2 tag slots of 2-8
    candidates, 0-30 char and 0-10 type tag n-grams each with one rel_position in [0, 3])."""
    L = 40
    cps = gen_codepoints(sample_sentences, L, TEXT_SEED ^ 0x77).reshape(sample_sentences, L)
    allk, allc, alln = [], [], []
    for n in (1, 2, 3):
        k, c = _ngram_counts(cps, n)
        allk.append(k)
        allc.append(c)
        alln.append(np.full(len(k), n, np.int8))
    keys = np.concatenate(allk)
    cnts = np.concatenate(allc)
    ns = np.concatenate(alln)
    order = np.lexsort((keys, ns, -cnts))[:n_patterns]
    keys, ns = keys[order], ns[order]
    ngram_bytes, rows = [], []
    for n in (1, 2, 3):
        sel = np.nonzero(ns == n)[0]
        width = 2 * window - n + 1
        wts = _weights(seed, len(sel), width, stream=10 + n)
        ks = keys[sel]
        cols = [((ks >> np.uint64(21 * (n - 1 - j))) & np.uint64(0x1FFFFF)).astype(np.uint32) for j in range(n)]
        cache = {}
        for i in range(len(sel)):
            g = b"".join(cache.setdefault(int(c[i]), _cp_to_utf8_bytes(c[i])) for c in cols)
            ngram_bytes.append(g)
            rows.append(wts[i])
    char_part = _encode_ngram_list(ngram_bytes, rows)
    # type n-grams: all sequences over 1..6 of length 1..3
    tgrams, trows = [], []
    for n in (1, 2, 3):
        width = 2 * window - n + 1
        cnt = 6 ** n
        wts = _weights(seed, cnt, width, stream=20 + n)
        for i in range(cnt):
            digs = []
            x = i
            for _ in range(n):
                digs.append(1 + x % 6)
                x //= 6
            tgrams.append(bytes(reversed(digs)))
            trows.append(wts[i])
    type_part = _encode_ngram_list(tgrams, trows)
    # dictionary
    dparts = [varint(dict_words)]
    if dict_words:
        u = _uniform(seed, dict_words, 31)
        lens = np.select([u < 0.02, u < 0.32, u < 0.57, u < 0.75, u < 0.92], [1, 2, 3, 4, 0], default=-1)
        r5 = (splitmix64(seed, dict_words, 32) % np.uint64(3)).astype(np.int64) + 5
        r8 = (splitmix64(seed, dict_words, 33) % np.uint64(9)).astype(np.int64) + 8
        lens = np.where(lens == 0, r5, np.where(lens == -1, r8, lens))
        # words are substrings of fresh synthetic text so that they actually occur
        wcps = gen_codepoints(dict_words, lens, TEXT_SEED ^ 0x99)
        wstart = np.zeros(dict_words + 1, np.int64)
        np.cumsum(lens, out=wstart[1:])
        bw = _weights(seed, 4, 3, stream=34)  # [bucket][L, I, R]
        bw = np.where(bw == 0, 1, bw)
        cache = {}
        seen = set()
        n_real = 0
        body = []
        for i in range(dict_words):
            cp = wcps[wstart[i]:wstart[i + 1]]
            wb = b"".join(cache.setdefault(int(c), _cp_to_utf8_bytes(c)) for c in cp)
            if wb in seen:
                continue
            seen.add(wb)
            n_real += 1
            ln = int(lens[i])
            b = min(ln, 4) - 1
            ws = [int(bw[b][0])] + [int(bw[b][1])] * (ln - 1) + [int(bw[b][2])]
            body.append(varint(len(wb)) + wb + varint(len(ws)) + b"".join(zigzag(x) for x in ws) + varint(0))
        dparts = [varint(n_real)] + body
    bias = int(splitmix64(seed, 1, 40)[0] % np.uint64(65535)) - 32767
    tparts = [varint(0)]
    if tag_models:
        r = splitmix64(seed, tag_models * 8, 50).reshape(tag_models, 8)
        tok_len = (r[:, 0] % np.uint64(4)).astype(np.int64) + 1
        tok_cps = gen_codepoints(tag_models, tok_len, TEXT_SEED ^ 0x55)
        tstart = np.zeros(tag_models + 1, np.int64)
        np.cumsum(tok_len, out=tstart[1:])
        n_char_ng = (r[:, 1] % np.uint64(31)).astype(np.int64)
        n_type_ng = (r[:, 2] % np.uint64(11)).astype(np.int64)
        ng_total = int(n_char_ng.sum())
        ng_len = (splitmix64(seed, ng_total, 51) % np.uint64(3)).astype(np.int64) + 1
        ng_cps = gen_codepoints(ng_total, ng_len, TEXT_SEED ^ 0x66)
        ng_start = np.zeros(ng_total + 1, np.int64)
        np.cumsum(ng_len, out=ng_start[1:])
        ng_rel = (splitmix64(seed, ng_total, 52) % np.uint64(window + 1)).astype(np.int64)
        tng_total = int(n_type_ng.sum())
        tng_r = splitmix64(seed, tng_total * 4, 53).reshape(tng_total, 4) if tng_total else np.zeros((0, 4), np.uint64)
        cache = {}
        seen = set()
        body = []
        ci = ti = 0
        wsrc = (splitmix64(seed, 1 << 16, 54) % np.uint64(2001)).astype(np.int64) - 1000
        wpos = 0

        def wvec(k):
            nonlocal wpos
            if wpos + k > len(wsrc):
                wpos = 0
            v = wsrc[wpos:wpos + k]
            wpos += k
            return v

        for i in range(tag_models):
            tok = b"".join(cache.setdefault(int(c), _cp_to_utf8_bytes(c)) for c in tok_cps[tstart[i]:tstart[i + 1]])
            nc, nt = int(n_char_ng[i]), int(n_type_ng[i])
            if tok in seen:
                ci += nc
                ti += nt
                continue
            seen.add(tok)
            nslots = int(r[i, 3] % np.uint64(2)) + 1
            cands = [int(r[i, 4 + s] % np.uint64(7)) + 2 for s in range(nslots)]
            slen = sum(cands)
            p = [varint(len(tok)), tok, varint(nslots)]
            for s_, c_ in enumerate(cands):
                p.append(varint(c_))
                for k in range(c_):
                    t = b"T%d_%d" % (s_, k)
                    p.append(varint(len(t)) + t)
            p.append(varint(nc))
            for j in range(nc):
                g = b"".join(cache.setdefault(int(c), _cp_to_utf8_bytes(c)) for c in ng_cps[ng_start[ci + j]:ng_start[ci + j + 1]])
                p.append(varint(len(g)) + g + varint(1) + bytes([int(ng_rel[ci + j])]) + varint(slen) +
                         b"".join(zigzag(int(x)) for x in wvec(slen)))
            ci += nc
            p.append(varint(nt))
            for j in range(nt):
                rr = tng_r[ti + j]
                ln = int(rr[0] % np.uint64(3)) + 1
                g = bytes(int(rr[1 + q] % np.uint64(6)) + 1 for q in range(ln))
                p.append(varint(len(g)) + g + varint(1) + bytes([int(rr[0] >> np.uint64(8)) % (window + 1)]) + varint(slen) +
                         b"".join(zigzag(int(x)) for x in wvec(slen)))
            ti += nt
            p.append(varint(slen) + b"".join(zigzag(int(x)) for x in wvec(slen)))
            body.append(b"".join(p))
        tparts = [varint(len(body))] + body
    out = b"".join([MODEL_MAGIC, char_part, type_part, b"".join(dparts), zigzag(bias), bytes([window]), bytes([window]),
                    b"".join(tparts)])
    return out
