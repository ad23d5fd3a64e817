"""Encode a Vaporetto model file from a plain-dict description.

Writer counterpart of the format read by `Model::read` (reference
vaporetto/src/model.rs:15,61-70,142-153): magic + bincode 2.0.1
`config::standard()` (varint ints, zigzag signed, length-prefixed Vec/String).
Used by tests to build the small models the reference's unit tests construct
in code (e.g. predictor.rs:749-838) and by the synthetic-model generator.
"""
from __future__ import annotations

import struct

MODEL_MAGIC = b"VaporettoTokenizer 0.5.0\n"


def varint(u: int) -> bytes:
    if u < 251:
        return bytes([u])
    if u < 1 << 16:
        return b"\xfb" + struct.pack("<H", u)
    if u < 1 << 32:
        return b"\xfc" + struct.pack("<I", u)
    return b"\xfd" + struct.pack("<Q", u)


def zigzag(i: int) -> bytes:
    return varint((i << 1) ^ (i >> 63) if i < 0 else i << 1)


def _bytes(b) -> bytes:
    if isinstance(b, str):
        b = b.encode("utf-8")
    b = bytes(b)
    return varint(len(b)) + b


def _vec_i32(v) -> bytes:
    return varint(len(v)) + b"".join(zigzag(int(x)) for x in v)


def _ngrams(lst) -> bytes:
    out = [varint(len(lst))]
    for ngram, weights in lst:
        out.append(_bytes(ngram))
        out.append(_vec_i32(weights))
    return b"".join(out)


def _tag_ngrams(lst) -> bytes:
    out = [varint(len(lst))]
    for ngram, tws in lst:
        out.append(_bytes(ngram))
        out.append(varint(len(tws)))
        for rel, weights in tws:
            out.append(bytes([rel]))
            out.append(_vec_i32(weights))
    return b"".join(out)


def encode_model(m: dict) -> bytes:
    """m keys: char_ngrams [(str, [w])], type_ngrams [(bytes, [w])], dict [(word, [w], comment)],
    bias, char_window, type_window, tag_models [ {token, tags [[str]], char_ngrams [(str, [(rel,[w])])],
    type_ngrams [(bytes, [(rel,[w])])], bias [w]} ]"""
    out = [MODEL_MAGIC]
    out.append(_ngrams(m.get("char_ngrams", [])))
    out.append(_ngrams(m.get("type_ngrams", [])))
    d = m.get("dict", [])
    out.append(varint(len(d)))
    for rec in d:
        word, weights = rec[0], rec[1]
        comment = rec[2] if len(rec) > 2 else ""
        out.append(_bytes(word))
        out.append(_vec_i32(weights))
        out.append(_bytes(comment))
    out.append(zigzag(int(m.get("bias", 0))))
    out.append(bytes([m.get("char_window", 0)]))
    out.append(bytes([m.get("type_window", 0)]))
    tms = m.get("tag_models", [])
    out.append(varint(len(tms)))
    for t in tms:
        out.append(_bytes(t["token"]))
        out.append(varint(len(t["tags"])))
        for cands in t["tags"]:
            out.append(varint(len(cands)))
            for c in cands:
                out.append(_bytes(c))
        out.append(_tag_ngrams(t.get("char_ngrams", [])))
        out.append(_tag_ngrams(t.get("type_ngrams", [])))
        out.append(_vec_i32(t.get("bias", [])))
    return b"".join(out)
