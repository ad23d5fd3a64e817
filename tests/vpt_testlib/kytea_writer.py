"""Writer of the KyTea binary model format as the reference reads it (vaporetto/src/kytea_model.rs:28-450), for tests:
builds model files from a small Python description so that the converter can be checked on more than the one bundled
fixture.  Layout (little endian): config line + fields + character map; the word-segmentation linear model with its
feature lookup (three tries of i16 vectors + four i16 vectors); per tag: global tags + a linear model; the word
dictionary (trie of tag entries); the sub-word dictionary (trie of probability entries)."""
import struct

import numpy as np


def u8(v):
    return struct.pack("<B", v)


def u16(v):
    return struct.pack("<H", v)


def u32(v):
    return struct.pack("<I", v)


def i32(v):
    return struct.pack("<i", v)


def f64(v):
    return struct.pack("<d", v)


class KyteaWriter:
    def __init__(self, chars: str):
        """chars: every character used by any string of the model (the character map; indices are 1-based)."""
        self.chars = list(dict.fromkeys(chars))
        self.index = {c: i + 1 for i, c in enumerate(self.chars)}

    def string(self, s: str) -> bytes:
        return u32(len(s)) + b"".join(u16(self.index[c]) for c in s)

    @staticmethod
    def vec_i16(v) -> bytes:
        return u32(len(v)) + b"".join(struct.pack("<h", int(x)) for x in v)

    def trie(self, items: dict, entry_bytes, n_dicts: int = 1, rng=None) -> bytes:
        """items: {string: entry}.  States in breadth-first order, transitions in random order (the reader sorts
        them), terminal states flagged `is_branch` with their entry first in `outputs`; some decoy outputs."""
        if not items:
            return u8(n_dicts) + u32(0)
        nodes = [{}]          # state -> {char: state}
        term = {}             # state -> entry index
        entries = []
        for key, ent in items.items():
            s = 0
            for c in key:
                if c not in nodes[s]:
                    nodes.append({})
                    nodes[s][c] = len(nodes) - 1
                s = nodes[s][c]
            term[s] = len(entries)
            entries.append(ent)
        out = [u8(n_dicts), u32(len(nodes))]
        for s, nxt in enumerate(nodes):
            gotos = list(nxt.items())
            if rng is not None:
                rng.shuffle(gotos)
            out.append(u32(0 if rng is None else int(rng.integers(0, len(nodes)))))  # failure link (ignored)
            out.append(u32(len(gotos)))
            for c, t in gotos:
                out.append(u16(self.index[c]) + u32(t))
            outputs = []
            if s in term:
                outputs.append(term[s])
                if rng is not None and rng.random() < 0.3:
                    outputs.append(int(rng.integers(0, len(entries))))    # a suffix match after the own entry
            elif rng is not None and rng.random() < 0.2 and entries:
                outputs.append(int(rng.integers(0, len(entries))))        # suffix match on a non-terminal state
            out.append(u32(len(outputs)) + b"".join(u32(o) for o in outputs))
            out.append(u8(1 if s in term else 0))
        out.append(u32(len(entries)))
        out.extend(entry_bytes(e) for e in entries)
        return b"".join(out)

    def linear_model(self, lm, rng=None) -> bytes:
        """lm: None, or dict(labels=[..], lookup=None | dict(chars={str: [i16]}, types={..}, selfs={..}, dict_vec=[..],
        biases=[..], tag_dict_vec=[..], tag_unk_vec=[..]))."""
        if lm is None:
            return u32(0)
        labels = lm.get("labels", [1, -1])
        out = [u32(len(labels)), u8(lm.get("solver", 1))] + [i32(x) for x in labels] + [u8(1), f64(lm.get("multiplier", 1.0))]
        lk = lm.get("lookup")
        if lk is None:
            out.append(u8(0))
        else:
            out.append(u8(1))
            for key in ("chars", "types", "selfs"):
                out.append(self.trie(lk.get(key, {}), self.vec_i16, rng=rng))
            for key in ("dict_vec", "biases", "tag_dict_vec", "tag_unk_vec"):
                out.append(self.vec_i16(lk.get(key, [])))
        return b"".join(out)

    def model(self, *, char_w, type_w, dict_n, wordseg, n_tags=0, global_tags=(), global_models=(), words=None,
              n_dicts=1, subwords=None, rng=None) -> bytes:
        """words: {word: dict(in_dict=bits, tags=[[(tag, in_dict_bits)...] per tag], models=[lm per tag])};
        subwords: {word: [[(tag, prob)...] per tag]}."""
        out = [b"KyTea test model\n", u8(1), u8(1 if n_tags else 0), u32(n_tags), u8(char_w), u8(3), u8(type_w), u8(3),
               u8(dict_n), u8(1), f64(1e-3), u8(1), "".join(self.chars).encode("utf-8") + b"\x00"]
        out.append(self.linear_model(wordseg, rng))
        for t in range(n_tags):
            tags = global_tags[t] if t < len(global_tags) else []
            out.append(u32(len(tags)) + b"".join(self.string(x) for x in tags))
            out.append(self.linear_model(global_models[t] if t < len(global_models) else None, rng))

        def word_entry(item):
            word, ent = item
            b = [self.string(word)]
            for t in range(n_tags):
                tl = ent.get("tags", [[]] * n_tags)[t]
                b.append(u32(len(tl)) + b"".join(self.string(tag) + u8(bits) for tag, bits in tl))
            b.append(u8(ent["in_dict"]))
            for t in range(n_tags):
                b.append(self.linear_model(ent.get("models", [None] * n_tags)[t], rng))
            return b"".join(b)

        out.append(self.trie({w: (w, e) for w, e in (words or {}).items()}, word_entry, n_dicts=n_dicts, rng=rng))

        def sub_entry(item):
            word, per_tag = item
            b = [self.string(word)]
            for t in range(n_tags):
                tl = per_tag[t] if t < len(per_tag) else []
                b.append(u32(len(tl)) + b"".join(self.string(tag) + f64(p) for tag, p in tl))
            return b"".join(b)

        out.append(self.trie({w: (w, e) for w, e in (subwords or {}).items()}, sub_entry, rng=rng))
        return b"".join(out)


def random_model(rng: np.random.Generator) -> bytes:
    """A random KyTea model file exercising every section the reader walks through."""
    alphabet = "あいう火星猫aB1。é\U00020000"
    types = "DRHTKO"
    w = KyteaWriter(alphabet + types + "\x04" + "名詞動tag")
    char_w, type_w, dict_n = (int(rng.integers(1, 4)) for _ in range(3))
    n_tags = int(rng.integers(0, 3))
    n_dicts = int(rng.integers(1, 4))

    def word(src, lo, hi):
        return "".join(rng.choice(list(src), size=int(rng.integers(lo, hi + 1))))

    def wv(n):
        return rng.integers(-30000, 30000, size=int(n)).tolist()

    def ngrams(src, window, extra=""):
        d = {}
        for _ in range(int(rng.integers(1, 25))):
            k = word(src + extra if rng.random() < 0.15 else src, 1, min(2 * window, 4))
            d[k] = wv(2 * window - len(k) + 1 + int(rng.integers(0, 3)))  # sometimes longer than needed: truncated
        return d

    def small_lm():
        if rng.random() < 0.4:
            return None
        lk = None if rng.random() < 0.3 else dict(chars=ngrams(alphabet, 2), types=ngrams(types, 2),
                                                  selfs={word(alphabet, 1, 2): wv(3)} if rng.random() < 0.5 else {},
                                                  dict_vec=wv(rng.integers(0, 5)), biases=wv(rng.integers(1, 4)),
                                                  tag_dict_vec=wv(rng.integers(0, 4)), tag_unk_vec=wv(rng.integers(0, 3)))
        return dict(labels=wv(rng.integers(1, 4)), lookup=lk)

    wordseg = dict(lookup=dict(chars=ngrams(alphabet, char_w), types=ngrams(types, type_w, extra="\x04"),
                               selfs={word(alphabet, 1, 3): wv(4)} if rng.random() < 0.5 else {},
                               dict_vec=wv(3 * dict_n * n_dicts), biases=wv(rng.integers(1, 3)),
                               tag_dict_vec=wv(rng.integers(0, 4)), tag_unk_vec=wv(rng.integers(0, 4))))
    words = {}
    for _ in range(int(rng.integers(0, 20))):
        words[word(alphabet, 1, 6)] = dict(
            in_dict=int(rng.integers(0, 1 << n_dicts)),
            tags=[[(word("名詞動tag", 1, 3), int(rng.integers(0, 4))) for _ in range(int(rng.integers(0, 3)))] for _ in range(n_tags)],
            models=[small_lm() for _ in range(n_tags)])
    subwords = {word(alphabet, 1, 2): [[(word("名詞動tag", 1, 2), float(rng.random())) for _ in range(int(rng.integers(0, 3)))]
                                      for _ in range(n_tags)] for _ in range(int(rng.integers(0, 4)))}
    return w.model(char_w=char_w, type_w=type_w, dict_n=dict_n, wordseg=wordseg, n_tags=n_tags,
                   global_tags=[[word("名詞動tag", 1, 3) for _ in range(int(rng.integers(0, 3)))] for _ in range(n_tags)],
                   global_models=[small_lm() for _ in range(n_tags)], words=words, n_dicts=n_dicts, subwords=subwords, rng=rng)
