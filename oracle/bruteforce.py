"""ORACLE #2 — TEST INFRASTRUCTURE ONLY.  Independent brute-force scorer.

Shares no code or data structure with oracle/vaporetto_oracle.cpp (no merger, no
automaton, no type table): the score of boundary i is computed straight from the
model description as

    bias + sum over every occurrence of every char n-gram / dictionary word / type n-gram
           of the un-merged weight that lands on boundary i.

This equals the reference's result because the reference's suffix-merged weights
(char_scorer.rs:50-78) reported once per end position for the longest match
(char_scorer/boundary_scorer.rs:95-113) are by construction the sum of the un-merged
weights of all patterns ending there, and i32 wrapping addition is associative and
commutative.  Positions follow predictor.rs:176-213:
an occurrence ending at char index `end` (exclusive) with offset `off` adds w[k] to
boundary end-1+off+k, dropped outside [0, n-1).
"""
from __future__ import annotations


def get_type(c: int) -> int:  # sentence.rs:50-67
    if 0x30 <= c <= 0x39 or 0xFF10 <= c <= 0xFF19:
        return 1
    if 0x41 <= c <= 0x5A or 0x61 <= c <= 0x7A or 0xFF21 <= c <= 0xFF3A or 0xFF41 <= c <= 0xFF5A:
        return 2
    if 0x3040 <= c <= 0x3096:
        return 3
    if 0x30A0 <= c <= 0x30FA or 0x30FC <= c <= 0x30FF or 0xFF66 <= c <= 0xFF9F:
        return 4
    if (0x3400 <= c <= 0x4DBF or 0x4E00 <= c <= 0x9FFF or 0xF900 <= c <= 0xFAFF or 0x20000 <= c <= 0x2A6DF
            or 0x2A700 <= c <= 0x2B73F or 0x2B740 <= c <= 0x2B81F or 0x2B820 <= c <= 0x2CEAF
            or 0x2F800 <= c <= 0x2FA1F):
        return 5
    return 6


def _wrap(x: int) -> int:
    x &= 0xFFFFFFFF
    return x - (1 << 32) if x >= 1 << 31 else x


def predict(model: dict, text: str):
    """model: the dict accepted by vpt_testlib.bincode_model.encode_model.  Returns (scores, boundaries)."""
    chars = list(text)
    n = len(chars)
    assert n >= 1
    types = [get_type(ord(c)) for c in chars]
    ys = [int(model.get("bias", 0))] * (n - 1)

    def add(end, off, ws):
        for k, w in enumerate(ws):
            i = end - 1 + off + k
            if 0 <= i < n - 1:
                ys[i] += w

    cw = model.get("char_window", 0)
    tw = model.get("type_window", 0)
    cng = model.get("char_ngrams", [])
    dic = model.get("dict", [])
    if cw > 0 and (cng or dic):
        for ngram, ws in cng:
            g = list(ngram)
            L = len(g)
            for end in range(L, n + 1):
                if chars[end - L:end] == g:
                    add(end, -cw, ws)
        for rec in dic:
            g = list(rec[0])
            L = len(g)
            for end in range(L, n + 1):
                if chars[end - L:end] == g:
                    add(end, -L, rec[1])
    tng = model.get("type_ngrams", [])
    if tw > 0 and tng:
        if tw <= 3:
            # cached table semantics (type_scorer/boundary_scorer_cache.rs:36-49,59-81): the 2W-type window
            # around boundary i, zero-padded outside the sentence; weight index = 2W - end_in_window.
            for i in range(n - 1):
                win = [(types[j] if 0 <= j < n else 0) for j in range(i - tw + 1, i + tw + 1)]
                for ngram, ws in tng:
                    g = list(ngram)
                    L = len(g)
                    for end in range(L, 2 * tw + 1):
                        if win[end - L:end] == g:
                            idx = 2 * tw - end
                            if idx < len(ws):
                                ys[i] += ws[idx]
        else:
            for ngram, ws in tng:
                g = list(ngram)
                L = len(g)
                for end in range(L, n + 1):
                    if types[end - L:end] == g:
                        add(end, -tw, ws)
    ys = [_wrap(y) for y in ys]
    return ys, [1 if y > 0 else 0 for y in ys]
