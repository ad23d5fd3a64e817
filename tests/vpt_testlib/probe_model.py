"""Host-side count of the node-table probes the streaming kernel (vaporetto_b200/csrc/fused_kernel.cuh) issues for a
given model and text: one probe of the 2-character node per character, then the 3-character node when the
2-character record's child mask has the bit of the preceding character (keys.hpp: child_bit), or the 1-character
node when the 2-character node does not exist.  Used by bench.py for the second roofline entry (random 32-byte
record loads against the measured line rate of the L1 tag stage).  Patterns longer than 3 characters (dictionary
words) add backward-walk probes that are not counted here."""
from __future__ import annotations

import struct

K_MUL = 0x9E3779B1
MASK_BITS = 19


def child_bit(c: int) -> int:
    return (((c * K_MUL) & 0xFFFFFFFF) * MASK_BITS) >> 32


def char_patterns(model_bytes: bytes):
    """n-gram strings and dictionary words of a serialized model (bincode standard config, model.rs:61-70)."""
    pos = 25
    data = model_bytes

    def varint():
        nonlocal pos
        b = data[pos]
        pos += 1
        if b < 251:
            return b
        if b == 251:
            v = struct.unpack_from("<H", data, pos)[0]
            pos += 2
            return v
        if b == 252:
            v = struct.unpack_from("<I", data, pos)[0]
            pos += 4
            return v
        v = struct.unpack_from("<Q", data, pos)[0]
        pos += 8
        return v

    def string():
        nonlocal pos
        n = varint()
        s = data[pos:pos + n]
        pos += n
        return s

    def skip_weights():
        for _ in range(varint()):
            varint()

    pats = []
    for _ in range(varint()):          # char n-grams
        pats.append(string().decode())
        skip_weights()
    for _ in range(varint()):          # type n-grams
        string()
        skip_weights()
    for _ in range(varint()):          # dictionary
        pats.append(string().decode())
        skip_weights()
        string()
    return pats


def probes_per_char(model_bytes: bytes, sentences) -> dict:
    nodes2, nodes3, nodes1 = set(), set(), set()
    for p in char_patterns(model_bytes):
        for i in range(len(p)):
            s = p[i:][-3:] if len(p) - i > 3 else p[i:]
            # every suffix of a pattern is a node; its last <= 3 characters are the shallow nodes on its path
            for k in (1, 2, 3):
                if len(s) >= k:
                    t = s[-k:]
                    (nodes1, nodes2, nodes3)[k - 1].add(t)
    masks = {}
    for t in nodes3:
        masks[t[1:]] = masks.get(t[1:], 0) | (1 << child_bit(ord(t[0])))
    n = first = second = 0
    for s in sentences:
        for p in range(len(s)):
            n += 1
            first += 1
            if p == 0:
                continue  # sentence start: the first probe is the 1-character node itself
            t2 = s[p - 1:p + 1]
            if t2 in nodes2:
                if p >= 2 and (masks.get(t2, 0) >> child_bit(ord(s[p - 2]))) & 1:
                    second += 1
            else:
                second += 1
    return {"chars": n, "probes": first + second, "per_char": (first + second) / max(n, 1)}
