"""vaporetto_b200 — Python mirror of the reference's `Model` / `Predictor` / `Sentence` API
(vaporetto/src/lib.rs:82-91) over the C ABI of libvaporetto_b200.so (include/vaporetto_b200.h).

The compute path is the CUDA library; there is no CPU fallback.  Importing this package works without a GPU
(so the ABI can be inspected), creating a `Predictor` does not.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

__all__ = ["Model", "Predictor", "Sentence", "VaporettoError", "CharacterBoundary", "CharacterType", "lib", "build",
           "BatchResult", "build_blob", "shard_by_bytes"]

_PKG = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("VPT_B200_LIBRARY") or os.path.join(_PKG, "libvaporetto_b200.so")  # (override: A/B builds)


def build(force: bool = False) -> str:
    """Compile the CUDA extension in-tree with nvcc for sm_100a (see csrc/Makefile)."""
    import subprocess
    csrc = os.path.join(_PKG, "csrc")
    if force and os.path.exists(_SO):
        os.remove(_SO)
    subprocess.check_call(["make", "-C", csrc, "-s"])
    return _SO


class VaporettoError(Exception):
    """Mirror of `VaporettoError` (vaporetto/src/errors.rs:15-38)."""

    KIND = {1: "InvalidModel", 2: "InvalidArgument", 3: "InvalidSentence", 4: "DecodeError", 5: "IOError",
            16: "CudaError", 17: "Unsupported", 18: "Internal"}

    def __init__(self, code: int, msg: str):
        super().__init__(msg)
        self.code = code
        self.kind = self.KIND.get(code, "Unknown")


class CharacterBoundary:  # sentence.rs:70-82
    NotWordBoundary = 0
    WordBoundary = 1
    Unknown = 2


class CharacterType:  # sentence.rs:9-29
    Digit, Roman, Hiragana, Katakana, Kanji, Other = 1, 2, 3, 4, 5, 6


class _Info(C.Structure):
    _fields_ = [("device", C.c_int32), ("predict_tags", C.c_int32), ("n_tags", C.c_int32), ("char_scorer", C.c_int32),
                ("type_scorer", C.c_int32), ("fast_path", C.c_int32), ("bias", C.c_int32), ("char_window", C.c_int32),
                ("type_window", C.c_int32), ("n_char_patterns", C.c_uint32), ("n_type_patterns", C.c_uint32),
                ("n_char_nodes", C.c_uint32), ("n_type_nodes", C.c_uint32), ("max_char_pattern_len", C.c_uint32),
                ("blob_bytes", C.c_uint64), ("kernel_launches_per_batch", C.c_int32)]


# every symbol include/vaporetto_b200.h declares: (name, restype, argtypes)
_P = C.c_void_p
ABI = [
    ("vpt_last_error", C.c_char_p, []),
    ("vpt_version", C.c_char_p, []),
    ("vpt_model_read", C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(_P), C.POINTER(C.c_size_t)]),
    ("vpt_model_free", None, [_P]),
    ("vpt_model_read_kytea", C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(_P)]),
    ("vpt_concat_grapheme_clusters", C.c_int, [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    ("vpt_split_linebreaks", C.c_int, [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    ("vpt_model_read_zstd", C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(_P)]),
    ("vpt_model_to_vec", C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_uint64)]),
    ("vpt_model_dictionary_len", C.c_uint64, [_P]),
    ("vpt_model_dictionary_get", C.c_int, [_P, C.c_uint64, C.POINTER(C.c_char_p), C.POINTER(_P), C.POINTER(C.c_uint64),
                                          C.POINTER(C.c_char_p)]),
    ("vpt_model_replace_dictionary", C.c_int, [_P, _P, _P, _P, _P, C.c_uint64]),
    ("vpt_predictor_new", C.c_int, [_P, C.c_int, C.c_int, C.POINTER(_P)]),
    ("vpt_predictor_free", None, [_P]),
    ("vpt_predictor_get_info", C.c_int, [_P, C.POINTER(_Info)]),
    ("vpt_blob_build", C.c_int, [_P, C.c_int, C.POINTER(_P), C.POINTER(C.c_uint64)]),
    ("vpt_blob_free", None, [_P]),
    ("vpt_predictor_blob_size", C.c_uint64, [_P]),
    ("vpt_predictor_blob_export", C.c_int, [_P, _P, C.c_uint64]),
    ("vpt_predictor_from_blob", C.c_int, [_P, C.c_uint64, C.c_int, C.POINTER(_P)]),
    ("vpt_predict_batch", C.c_int, [_P, _P, _P, C.c_size_t, _P, _P, C.c_size_t, _P, _P, _P, _P, C.c_size_t, _P,
                                    C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    ("vpt_workspace_size", C.c_uint64, [C.c_size_t]),
    ("vpt_predict_batch_dev", C.c_int, [_P, _P, _P, C.c_size_t, _P, C.c_uint64, _P, _P, _P, _P, _P, _P, _P, _P]),
    ("vpt_predict_batch_dev_profiled", C.c_int, [_P, _P, _P, C.c_size_t, _P, C.c_uint64, _P, _P, _P, _P, _P, _P, _P, _P,
                                                 _P]),
    ("vpt_predict", C.c_int, [_P, C.c_char_p, C.c_size_t, _P, _P, C.c_size_t, _P, _P, C.c_size_t, C.POINTER(C.c_uint64)]),
    ("vpt_fill_tags", C.c_int, [_P, C.c_char_p, C.c_size_t, _P, _P, _P, _P, _P, _P, C.c_size_t]),
    ("vpt_tag_string", C.c_char_p, [_P, C.c_uint32, C.c_uint32, C.c_uint32]),
    ("vpt_tag_n_candidates", C.c_uint32, [_P, C.c_uint32, C.c_uint32]),
    ("vpt_tag_score_len", C.c_uint32, [_P, C.c_uint32]),
    ("vpt_tag_n_tokens", C.c_uint32, [_P]),
    ("vpt_char_types", C.c_int, [C.c_char_p, C.c_size_t, _P, C.c_size_t, C.POINTER(C.c_uint64)]),
    ("vpt_write_tokenized_text", C.c_int, [_P, C.c_char_p, C.c_size_t, _P, _P, _P, _P, C.c_size_t,
                                           C.POINTER(C.c_uint64)]),
    ("vpt_tokenize_lines", C.c_int, [_P, _P, C.c_size_t, C.c_int, C.c_uint32, _P, C.c_size_t, C.POINTER(C.c_uint64),
                                     C.POINTER(C.c_uint64)]),
    ("vpt_kytea_fullwidth", C.c_uint32, [C.c_uint32]),
    ("vpt_device_pci_bus_id", C.c_int, [C.c_int, C.c_char_p, C.c_size_t]),
    ("vpt_predict_tags_batch_dev", C.c_int, [_P, _P, _P, C.c_size_t, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    ("vpt_predict_batch_tags", C.c_int, [_P, _P, _P, C.c_size_t, _P, _P, C.c_size_t, _P, _P, _P, _P, C.c_size_t, _P,
                                         C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    ("vpt_predict_batch_compact", C.c_int, [_P, _P, _P, C.c_size_t, _P, C.c_size_t, _P, _P, _P, _P, _P, C.c_size_t,
                                            C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    ("vpt_unpack_boundaries", C.c_int, [_P, C.c_uint64, C.c_uint64, _P]),
    ("vpt_tokenize_lines_tags", C.c_int, [_P, _P, C.c_size_t, C.c_int, C.c_uint32, _P, C.c_size_t,
                                          C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
]

_lib = None


def lib():
    """Loads libvaporetto_b200.so; fails loudly if the extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise ImportError(f"{_SO} is missing: the CUDA extension must be built first "
                              f"(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
        L = C.CDLL(_SO)
        for name, res, args in ABI:
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _check(rc: int):
    if rc != 0:
        raise VaporettoError(rc, lib().vpt_last_error().decode("utf-8", "replace"))


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


class Model:
    """`vaporetto::Model` (model.rs:58-168): an on-disk model image."""

    def __init__(self, handle, consumed: int):
        self._h = handle
        self.consumed = consumed

    @classmethod
    def read(cls, src) -> "Model":
        """`Model::read` (model.rs:142-153): file-like object or bytes with the raw (un-zstd'd) model."""
        data = src if isinstance(src, (bytes, bytearray, memoryview)) else src.read()
        return cls.read_slice(bytes(data))[0]

    @classmethod
    def read_slice(cls, data: bytes):
        """`Model::read_slice` (model.rs:127-134): returns (model, remaining bytes)."""
        h = _P()
        used = C.c_size_t()
        _check(lib().vpt_model_read(data, len(data), C.byref(h), C.byref(used)))
        return cls(h, used.value), data[used.value:]

    @classmethod
    def read_zstd(cls, src) -> "Model":
        """`Model::read(&mut zstd::Decoder::new(file)?)` (predict/src/main.rs:110-111): a *.model.zst image, decoded by
        the library (libzstd.so.1); a raw model image is accepted as well."""
        data = src if isinstance(src, (bytes, bytearray, memoryview)) else src.read()
        data = bytes(data)
        h = _P()
        _check(lib().vpt_model_read_zstd(data, len(data), C.byref(h)))
        return cls(h, len(data))

    @classmethod
    def read_kytea(cls, src) -> "Model":
        """`KyteaModel::read` + `Model::try_from` (kytea_model.rs:423-550): converts a KyTea binary model."""
        data = src if isinstance(src, (bytes, bytearray, memoryview)) else src.read()
        data = bytes(data)
        h = _P()
        _check(lib().vpt_model_read_kytea(data, len(data), C.byref(h)))
        return cls(h, len(data))

    def to_vec(self) -> bytes:
        """`Model::to_vec` (model.rs:99-104): the model file image."""
        if self._h is None:
            raise VaporettoError(2, "InvalidArgumentError: model: already consumed by Predictor::new")
        out = _P()
        n = C.c_uint64()
        _check(lib().vpt_model_to_vec(self._h, C.byref(out), C.byref(n)))
        try:
            return C.string_at(out, n.value)
        finally:
            lib().vpt_blob_free(out)

    def dictionary(self):
        """`Model::dictionary` (model.rs:155-158): [(word, weights, comment)]."""
        out = []
        for i in range(lib().vpt_model_dictionary_len(self._h)):
            w, c = C.c_char_p(), C.c_char_p()
            p, n = _P(), C.c_uint64()
            _check(lib().vpt_model_dictionary_get(self._h, i, C.byref(w), C.byref(p), C.byref(n), C.byref(c)))
            weights = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int32)), shape=(n.value,)).tolist() if n.value else []
            out.append((w.value.decode("utf-8"), weights, c.value.decode("utf-8")))
        return out

    def replace_dictionary(self, records) -> None:
        """`Model::replace_dictionary` (model.rs:160-163); records: [(word, weights, comment)] checked like
        `WordWeightRecord::new` (dict_model.rs:39-50)."""
        n = len(records)
        words = (C.c_char_p * max(n, 1))(*[r[0].encode("utf-8") for r in records])
        comments = (C.c_char_p * max(n, 1))(*[r[2].encode("utf-8") for r in records])
        arrays = [np.ascontiguousarray(r[1], np.int32) for r in records]
        wptr = (_P * max(n, 1))(*[a.ctypes.data for a in arrays])
        lens = (C.c_uint64 * max(n, 1))(*[a.size for a in arrays])
        _check(lib().vpt_model_replace_dictionary(self._h, words, wptr, lens, comments, n))

    def _take(self):
        h, self._h = self._h, None
        if h is None:
            raise VaporettoError(2, "InvalidArgumentError: model: already consumed by Predictor::new")
        return h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:  # (module globals are already gone when the interpreter shuts down)
            _lib.vpt_model_free(h)


def build_blob(model: Model, predict_tags: bool = False) -> np.ndarray:
    """Host-only build of the flat device model (vpt_blob_build); consumes `model`.  The blob can be broadcast
    as bytes and turned into a predictor on every rank with Predictor.from_blob."""
    out = _P()
    n = C.c_uint64()
    _check(lib().vpt_blob_build(model._take(), int(predict_tags), C.byref(out), C.byref(n)))
    try:
        return np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_uint8)), shape=(n.value,)).copy()
    finally:
        lib().vpt_blob_free(out)


def shard_by_bytes(offsets, rank: int, world: int):
    """Contiguous sentence range [lo, hi) of `rank` when a batch is split over `world` ranks balanced by bytes
    (SURVEY.md §8e).  Every sentence belongs to exactly one rank."""
    off = np.asarray(offsets, np.uint64)
    n = off.size - 1
    total = int(off[-1] - off[0])
    cuts = [int(np.searchsorted(off, int(off[0]) + total * r // world, side="left")) for r in range(world + 1)]
    cuts[0], cuts[-1] = 0, n
    cuts = [min(max(c, 0), n) for c in cuts]
    for i in range(1, world + 1):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return cuts[rank], cuts[rank + 1]


def shard_lines(data, rank: int, world: int):
    """Byte range [lo, hi) of `rank` when a buffer of lines (the input of Predictor.tokenize_lines) is split over
    `world` ranks: cuts are placed after the first '\n' at or after every r/world-th byte, so every line belongs to
    exactly one rank and the concatenation of the ranks' outputs equals the single-process output."""
    t = np.frombuffer(data, np.uint8) if isinstance(data, (bytes, bytearray)) else np.ascontiguousarray(data, np.uint8)
    n = t.size
    cuts = [0]
    for r in range(1, world):
        p = max(n * r // world, cuts[-1])
        nl = np.flatnonzero(t[p:] == 10)
        cuts.append(p + int(nl[0]) + 1 if p < n and nl.size else n)
    cuts.append(n)
    return cuts[rank], cuts[rank + 1]


class BatchResult:
    """Outputs of a batched predict: flat arrays + per-sentence offsets."""

    def __init__(self, scores, boundaries, bound_offsets, status, char_states=None, type_states=None,
                 char_offsets=None):
        self.scores = scores
        self.boundaries = boundaries
        self.bound_offsets = bound_offsets
        self.status = status
        self.char_states = char_states
        self.type_states = type_states
        self.char_offsets = char_offsets

    def sentence_scores(self, i: int) -> np.ndarray:
        return self.scores[int(self.bound_offsets[i]):int(self.bound_offsets[i + 1])]

    def sentence_boundaries(self, i: int) -> np.ndarray:
        return self.boundaries[int(self.bound_offsets[i]):int(self.bound_offsets[i + 1])]


class Predictor:
    """`vaporetto::Predictor` (predictor.rs:434-665) resident on one CUDA device."""

    def __init__(self, model: Model, predict_tags: bool = False, device: int = 0):
        """`Predictor::new(model, predict_tags)` — consumes `model` (predictor.rs:450)."""
        h = _P()
        _check(lib().vpt_predictor_new(model._take(), int(predict_tags), device, C.byref(h)))
        self._h = h
        self._load_info()

    @classmethod
    def from_blob(cls, blob, device: int = 0) -> "Predictor":
        self = cls.__new__(cls)
        b = np.ascontiguousarray(np.frombuffer(blob, dtype=np.uint8))
        h = _P()
        _check(lib().vpt_predictor_from_blob(b.ctypes.data, b.size, device, C.byref(h)))
        self._h = h
        self._load_info()
        return self

    def _load_info(self):
        info = _Info()
        _check(lib().vpt_predictor_get_info(self._h, C.byref(info)))
        self.info = {k: getattr(info, k) for k, _ in _Info._fields_}
        self.n_tags = info.n_tags
        self.predict_tags = bool(info.predict_tags)

    def export_blob(self) -> np.ndarray:
        n = lib().vpt_predictor_blob_size(self._h)
        out = np.empty(n, np.uint8)
        _check(lib().vpt_predictor_blob_export(self._h, out.ctypes.data, n))
        return out

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:  # (module globals are already gone when the interpreter shuts down)
            _lib.vpt_predictor_free(h)

    # -- predict ------------------------------------------------------------------------------------
    def predict(self, sentence: "Sentence") -> None:
        """`Predictor::predict(&self, &mut Sentence)` (predictor.rs:518-543)."""
        b = sentence._bytes
        n = sentence._n
        scores = np.zeros(max(n - 1, 1), np.int32)
        bounds = np.zeros(max(n - 1, 1), np.uint8)
        want = self.info["char_scorer"] == 2 or self.info["type_scorer"] == 3
        cs = np.full(n, 0xFFFFFFFF, np.uint32) if want else None
        ts = np.full(n, 0xFFFFFFFF, np.uint32) if want else None
        nch = C.c_uint64()
        _check(lib().vpt_predict(self._h, b, len(b), scores.ctypes.data, bounds.ctypes.data, max(n - 1, 1), _ptr(cs),
                                 _ptr(ts), n, C.byref(nch)))
        assert nch.value == n
        sentence._scores = scores[: n - 1]
        sentence._boundaries = bounds[: n - 1].copy()
        sentence._char_states = cs
        sentence._type_states = ts
        sentence._predictor = self
        sentence._tags = None

    def predict_batch(self, text, offsets, want_scores: bool = True, want_states: bool = False,
                      out: Optional[BatchResult] = None) -> BatchResult:
        """Batched predict over host buffers (vpt_predict_batch).  text: uint8 array / bytes,
        offsets: uint64 [n+1].  Outputs are sized from the text length unless `out` supplies buffers."""
        t = np.frombuffer(text, np.uint8) if isinstance(text, (bytes, bytearray)) else np.ascontiguousarray(text, np.uint8)
        off = np.ascontiguousarray(offsets, np.uint64)
        n = off.size - 1
        nbytes = int(off[-1] - off[0]) if n > 0 else 0
        if out is None:
            cap = max(nbytes, 1)
            out = BatchResult(np.empty(cap, np.int32) if want_scores else None, np.empty(cap, np.uint8),
                              np.empty(n + 1, np.uint64), np.empty(max(n, 1), np.int32),
                              np.empty(cap, np.uint32) if want_states else None,
                              np.empty(cap, np.uint32) if want_states else None,
                              np.empty(n + 1, np.uint64))
        nb = C.c_uint64()
        nc = C.c_uint64()
        _check(lib().vpt_predict_batch(self._h, t.ctypes.data, off.ctypes.data, n, _ptr(out.scores),
                                       out.boundaries.ctypes.data, out.boundaries.size, out.bound_offsets.ctypes.data,
                                       _ptr(out.status), _ptr(out.char_states), _ptr(out.type_states),
                                       0 if out.char_states is None else out.char_states.size,
                                       _ptr(out.char_offsets), C.byref(nb), C.byref(nc)))
        res = BatchResult(None if out.scores is None else out.scores[: nb.value], out.boundaries[: nb.value],
                          out.bound_offsets, out.status[:n],
                          None if out.char_states is None else out.char_states[: nc.value],
                          None if out.type_states is None else out.type_states[: nc.value], out.char_offsets)
        res.n_boundaries = nb.value
        res.n_chars = nc.value
        return res

    def predict_batch_tags(self, text, offsets, want_scores: bool = True):
        """predict + predict_tags for a batch with tag prediction on the device (vpt_predict_batch_tags).  Returns
        (BatchResult, tag_token [chars] int32, tag_cand [chars, n_tags] int32, n_unserved): the arrays `fill_tags`
        computes per sentence, indexed by BatchResult.char_offsets."""
        t = np.frombuffer(text, np.uint8) if isinstance(text, (bytes, bytearray)) else np.ascontiguousarray(text, np.uint8)
        off = np.ascontiguousarray(offsets, np.uint64)
        n = off.size - 1
        cap = max(int(off[-1] - off[0]) if n > 0 else 0, 1)
        nt = max(self.n_tags, 1)
        scores = np.empty(cap, np.int32) if want_scores else None
        bounds = np.empty(cap, np.uint8)
        boff = np.empty(n + 1, np.uint64)
        coff = np.empty(n + 1, np.uint64)
        status = np.empty(max(n, 1), np.int32)
        tok = np.empty(cap, np.int32)
        cand = np.empty(cap * nt, np.int32)
        nb, nc, nu = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(lib().vpt_predict_batch_tags(self._h, t.ctypes.data, off.ctypes.data, n, _ptr(scores), bounds.ctypes.data, cap,
                                            boff.ctypes.data, status.ctypes.data, tok.ctypes.data, cand.ctypes.data, cap,
                                            coff.ctypes.data, C.byref(nb), C.byref(nc), C.byref(nu)))
        res = BatchResult(None if scores is None else scores[: nb.value], bounds[: nb.value], boff, status[:n], None, None, coff)
        return res, tok[: nc.value], cand[: nc.value * nt].reshape(-1, nt), int(nu.value)

    def tag_string(self, token_id: int, slot: int, cand: int) -> Optional[str]:
        """Tag string of (token id, tag slot, candidate) as the token records of predict_batch_compact name it."""
        v = lib().vpt_tag_string(self._h, int(token_id), int(slot), int(cand))
        return None if v is None else v.decode("utf-8")

    def predict_batch_compact(self, text, offsets, tags: bool = False) -> "CompactResult":
        """predict (+ predict_tags) for a batch with compact results (vpt_predict_batch_compact): one bit per boundary, one
        record per token; see CompactResult."""
        t = np.frombuffer(text, np.uint8) if isinstance(text, (bytes, bytearray)) else np.ascontiguousarray(text, np.uint8)
        off = np.ascontiguousarray(offsets, np.uint64)
        n = off.size - 1
        cap = max(int(off[-1] - off[0]) if n > 0 else 0, 1)
        nt = self.n_tags if tags else 0
        bits = np.zeros((cap + 31) // 32 + 1, np.uint32)
        n_chars = np.zeros(max(n, 1), np.uint32)
        status = np.zeros(max(n, 1), np.uint8)
        n_tokens = np.zeros(max(n, 1), np.uint32)
        tok = np.empty(cap, np.int32) if tags else None
        cand = np.empty(cap * max(nt, 1), np.uint8) if tags else None
        nb, ntok, nu = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(lib().vpt_predict_batch_compact(self._h, t.ctypes.data, off.ctypes.data, n, bits.ctypes.data, bits.size,
                                               n_chars.ctypes.data, status.ctypes.data, n_tokens.ctypes.data, _ptr(tok), _ptr(cand),
                                               cap if tags else 0, C.byref(nb), C.byref(ntok), C.byref(nu)))
        return CompactResult(bits[: (nb.value + 31) // 32], int(nb.value), n_chars[:n], status[:n], n_tokens[:n],
                             None if tok is None else tok[: ntok.value],
                             None if cand is None else cand[: ntok.value * max(nt, 1)].reshape(-1, max(nt, 1)), int(nu.value))

    def tokenize_lines(self, data, out: Optional[np.ndarray] = None, no_norm: bool = False, wsconst: str = "",
                       predict_tags: bool = False):
        """The reference CLI's `predict` loop (predict/src/main.rs:126-181) over a whole buffer of raw bytes
        (vpt_tokenize_lines): lines are split, scored (on KyteaFullwidthFilter(line) unless no_norm) and written
        out as space-separated tokens on the device; `wsconst`: letters of the CLI's --wsconst options ("D", "DR", ...:
        KyteaWsConstFilter; "G": ConcatGraphemeClustersFilter); `predict_tags`: the CLI's --predict-tags
        (vpt_tokenize_lines_tags).  Returns (uint8 view of the output lines, number of lines)."""
        mask = 0
        for ch in wsconst:
            if ch not in "DRHTKOG":
                raise VaporettoError(2, "InvalidArgumentError: wsconst: one of D, R, H, T, K, O, G")
            mask |= 1 << ("DRHTKOG".index(ch) + 1)
        t = np.frombuffer(data, np.uint8) if isinstance(data, (bytes, bytearray)) else np.ascontiguousarray(data, np.uint8)
        if out is None:
            out = np.empty((3 + (16 if predict_tags else 0)) * t.size + int(np.count_nonzero(t == 10)) + 16, np.uint8)
        n = C.c_uint64()
        nl = C.c_uint64()
        fn = lib().vpt_tokenize_lines_tags if predict_tags else lib().vpt_tokenize_lines
        for _ in range(2):
            rc = fn(self._h, t.ctypes.data, t.size, int(no_norm), mask, out.ctypes.data, out.size, C.byref(n), C.byref(nl))
            if rc == 2 and predict_tags and n.value > out.size:
                out = np.empty(n.value + 16, np.uint8)   # long tag strings: the call reported the size it needs
                continue
            break
        _check(rc)
        return out[: n.value], int(nl.value)


class CompactResult:
    """Result of Predictor.predict_batch_compact: `boundary_bits` (uint32 words of the batch's boundary bit stream),
    `n_chars` / `status` / `n_tokens` per sentence, `token_ids` [tokens] and `token_cands` [tokens, n_tags] (uint8, 255 =
    none) when tags were requested.  Sentence s owns the bits [bit_offsets[s], bit_offsets[s + 1]) and the token records
    [token_offsets[s], token_offsets[s + 1])."""

    def __init__(self, bits, n_boundaries, n_chars, status, n_tokens, token_ids, token_cands, n_unserved):
        self.boundary_bits, self.n_boundaries = bits, n_boundaries
        self.n_chars, self.status, self.n_tokens = n_chars, status, n_tokens
        self.token_ids, self.token_cands, self.n_unserved = token_ids, token_cands, n_unserved
        nb = np.where(n_chars > 0, n_chars.astype(np.int64) - 1, 0)
        self.bit_offsets = np.concatenate(([0], np.cumsum(nb))).astype(np.uint64)
        self.token_offsets = np.concatenate(([0], np.cumsum(n_tokens.astype(np.int64)))).astype(np.uint64)

    def boundaries(self, s: Optional[int] = None) -> np.ndarray:
        """Boundaries as bytes (0 / 1): of sentence `s`, or of the whole batch (vpt_unpack_boundaries)."""
        lo, hi = (0, self.n_boundaries) if s is None else (int(self.bit_offsets[s]), int(self.bit_offsets[s + 1]))
        out = np.empty(hi - lo, np.uint8)
        _check(lib().vpt_unpack_boundaries(self.boundary_bits.ctypes.data, lo, hi - lo, out.ctypes.data))
        return out


class Token:
    """`vaporetto::Token` (sentence.rs:1195-1258)."""

    def __init__(self, sentence: "Sentence", start: int, end: int):
        self._s, self._start, self._end = sentence, start, end

    def surface(self) -> str:
        p = self._s._pos
        return self._s._bytes[p[self._start]:p[self._end]].decode("utf-8")

    def start(self) -> int:
        return self._start

    def end(self) -> int:
        return self._end

    def tags(self) -> List[Optional[str]]:
        k = self._s.n_tags()
        return self._s.tags()[(self._end - 1) * k:self._end * k]


class Sentence:
    """`vaporetto::Sentence` (sentence.rs:85-1193), raw-text subset used around `predict`."""

    def __init__(self, text: str):
        self._set(text)

    @classmethod
    def from_raw(cls, text: str) -> "Sentence":
        """`Sentence::from_raw` (sentence.rs:217-247)."""
        return cls(text)

    def update_raw(self, text: str) -> None:
        """`Sentence::update_raw` (sentence.rs:264-283); on error the sentence becomes " "."""
        try:
            self._set(text)
        except VaporettoError:
            self._set(" ")
            raise

    def _set(self, text: str):
        b = text.encode("utf-8") if isinstance(text, str) else bytes(text)
        types = np.zeros(max(len(b), 1), np.uint8)
        n = C.c_uint64()
        _check(lib().vpt_char_types(b, len(b), types.ctypes.data, types.size, C.byref(n)))
        self._bytes = b
        self._n = n.value
        self._types = types[: self._n].copy()
        starts = [i for i, c in enumerate(b) if (c & 0xC0) != 0x80]
        self._pos = starts + [len(b)]
        self._boundaries = np.full(self._n - 1, CharacterBoundary.Unknown, np.uint8)
        self._scores = np.zeros(0, np.int32)
        self._char_states = None
        self._type_states = None
        self._predictor = None
        self._tags = None
        self._tag_token = None
        self._tag_cand = None

    def as_raw_text(self) -> str:
        return self._bytes.decode("utf-8")

    def char_types(self) -> np.ndarray:
        return self._types

    def boundaries(self) -> np.ndarray:
        return self._boundaries

    def boundaries_mut(self) -> np.ndarray:
        return self._boundaries

    def split_linebreaks(self) -> None:
        """`SplitLinebreaksFilter::filter(&mut sentence)` (vaporetto_rules, split_linebreaks.rs:9-37)."""
        b = self.as_raw_text().encode("utf-8")
        bd = np.ascontiguousarray(self._boundaries, np.uint8)
        _check(lib().vpt_split_linebreaks(b, len(b), bd.ctypes.data, bd.size))
        self._boundaries[:] = bd

    def concat_grapheme_clusters(self) -> None:
        """`ConcatGraphemeClustersFilter::filter(&mut sentence)` (vaporetto_rules, concat_grapheme_clusters.rs:10-35)."""
        b = self.as_raw_text().encode("utf-8")
        bd = np.ascontiguousarray(self._boundaries, np.uint8)
        _check(lib().vpt_concat_grapheme_clusters(b, len(b), bd.ctypes.data, bd.size))
        self._boundaries[:] = bd

    def boundary_scores(self) -> np.ndarray:
        """`Sentence::boundary_scores` (sentence.rs:1040-1046)."""
        return self._scores

    def n_tags(self) -> int:
        return 0 if self._tags is None else self._predictor.n_tags

    def fill_tags(self) -> None:
        """`Sentence::fill_tags` (sentence.rs:1144-1148) -> `Predictor::predict_tags` (predictor.rs:546-637)."""
        p = self._predictor
        if p is None:
            return
        k = p.n_tags
        tt = np.full(self._n, -1, np.int32)
        tc = np.full(max(self._n * k, 1), -1, np.int32)
        _check(lib().vpt_fill_tags(p._h, self._bytes, len(self._bytes), self._boundaries.ctypes.data,
                                   _ptr(self._char_states), _ptr(self._type_states), tt.ctypes.data, tc.ctypes.data,
                                   None, 0))
        self._tag_token, self._tag_cand = tt, tc
        tags: List[Optional[str]] = []
        for i in range(self._n):
            for s in range(k):
                c = tc[i * k + s]
                if tt[i] < 0 or c < 0:
                    tags.append(None)
                else:
                    tags.append(lib().vpt_tag_string(p._h, int(tt[i]), s, int(c)).decode("utf-8"))
        self._tags = tags

    def tags(self) -> List[Optional[str]]:
        return [] if self._tags is None else self._tags

    def iter_tokens(self):
        """`Sentence::iter_tokens` (sentence.rs:819; TokenIterator :1273-1299)."""
        start, skip = 0, False
        for i, b in enumerate(self._boundaries):
            if b == CharacterBoundary.WordBoundary:
                if not skip:
                    yield Token(self, start, i + 1)
                skip = False
                start = i + 1
            elif b == CharacterBoundary.Unknown:
                skip = True
        if not skip:
            yield Token(self, start, self._n)

    def write_tokenized_text(self) -> str:
        """`Sentence::write_tokenized_text` (sentence.rs:850-886)."""
        p = self._predictor
        cap = 2 * len(self._bytes) + 64
        ln = C.c_uint64()
        have = self._tags is not None
        for _ in range(2):
            # the call reports the full length even when it truncated: retry once with the exact size
            buf = C.create_string_buffer(cap)
            _check(lib().vpt_write_tokenized_text(p._h if p else None, self._bytes, len(self._bytes),
                                                  self._boundaries.ctypes.data,
                                                  self._tag_token.ctypes.data if have else None,
                                                  self._tag_cand.ctypes.data if have else None, buf, cap, C.byref(ln)))
            if ln.value < cap:
                break
            cap = ln.value + 1
        if ln.value >= cap:
            raise VaporettoError(18, "write_tokenized_text: length changed between calls")
        return buf.raw[: ln.value].decode("utf-8")
