"""ctypes binding of the CPU oracle (oracle/libvaporetto_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never by the product.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
_ORA_DIR = os.path.join(_ROOT, "oracle")
_SO = os.path.join(_ORA_DIR, "libvaporetto_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_ORA_DIR, "vaporetto_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _ORA_DIR, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.ora_last_error.restype = C.c_char_p
        L.ora_model_read.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.ora_model_free.argtypes = [C.c_void_p]
        L.ora_model_from_kytea.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p)]
        L.ora_model_to_vec.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.ora_model_to_vec.restype = C.c_long
        L.ora_predictor_new.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.ora_predictor_free.argtypes = [C.c_void_p]
        L.ora_predictor_n_tags.argtypes = [C.c_void_p]
        L.ora_type_variant.argtypes = [C.c_void_p]
        L.ora_dump_patterns.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_size_t]
        L.ora_dump_patterns.restype = C.c_long
        L.ora_predict.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ora_predict.restype = C.c_long
        L.ora_char_types.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p]
        L.ora_char_types.restype = C.c_long
        L.ora_tokenize.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_int, C.c_char_p, C.c_size_t]
        L.ora_tokenize.restype = C.c_long
        L.ora_tokenize_lines.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_int, C.c_uint32, C.c_char_p,
                                         C.c_size_t, C.POINTER(C.c_uint64)]
        L.ora_tokenize_lines.restype = C.c_long
        L.ora_tokenize_lines_tags.argtypes = L.ora_tokenize_lines.argtypes
        L.ora_tokenize_lines_tags.restype = C.c_long
        L.ora_grapheme_lengths.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.ora_grapheme_lengths.restype = C.c_long
        L.ora_kytea_fullwidth.argtypes = [C.c_uint32]
        L.ora_kytea_fullwidth.restype = C.c_uint32
        L.ora_predict_tags.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.ora_predict_tags.restype = C.c_long
        L.ora_add_tag_scores.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_size_t, C.c_uint32, C.c_size_t,
                                         C.c_void_p, C.c_size_t]
        L.ora_predict_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_int]
        L.ora_count_chars.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        _lib = L
    return _lib


class OracleError(Exception):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


def _err(code):
    return OracleError(code, lib().ora_last_error().decode("utf-8", "replace"))


class OraclePredictor:
    """Model::read + Predictor::new + predict, on the CPU oracle."""

    def __init__(self, model_bytes: bytes, predict_tags: bool = False, states_only: bool = False):
        """states_only: tag predictor without the build-time merge of the tag weights (pattern ids, scores and
        boundaries are the reference's; predict_tags is not usable) -- for full-size tag models."""
        L = lib()
        L.ora_set_states_only(1 if states_only else 0)
        m = C.c_void_p()
        consumed = C.c_size_t()
        rc = L.ora_model_read(model_bytes, len(model_bytes), C.byref(m), C.byref(consumed))
        if rc:
            raise _err(rc)
        self.consumed = consumed.value
        p = C.c_void_p()
        rc = L.ora_predictor_new(m, int(predict_tags), C.byref(p))
        L.ora_set_states_only(0)
        L.ora_model_free(m)
        if rc:
            raise _err(rc)
        self._p = p
        self.n_tags = L.ora_predictor_n_tags(p)
        self.type_variant = L.ora_type_variant(p)

    def __del__(self):
        if getattr(self, "_p", None):
            lib().ora_predictor_free(self._p)
            self._p = None

    def predict(self, text, states: bool = False):
        b = text.encode("utf-8") if isinstance(text, str) else bytes(text)
        cap = max(len(b), 1)
        scores = np.zeros(cap, np.int32)
        bounds = np.zeros(cap, np.uint8)
        cs = np.zeros(cap, np.uint32)
        ts = np.zeros(cap, np.uint32)
        n = lib().ora_predict(self._p, b, len(b), scores.ctypes.data, bounds.ctypes.data,
                              cs.ctypes.data if states else None, ts.ctypes.data if states else None)
        if n < 0:
            raise _err(-n)
        if states:
            return scores[: n - 1].copy(), bounds[: n - 1].copy(), cs[:n].copy(), ts[:n].copy()
        return scores[: n - 1].copy(), bounds[: n - 1].copy()

    def tokenize(self, text: str, fill_tags: bool = False) -> str:
        b = text.encode("utf-8")
        cap = 16 * len(b) + 4096
        buf = C.create_string_buffer(cap)
        n = lib().ora_tokenize(self._p, b, len(b), int(fill_tags), buf, cap)
        if n < 0:
            raise _err(-n)
        return buf.raw[:n].decode("utf-8")

    def tokenize_lines(self, data: bytes, no_norm: bool = False, wsconst: str = "", predict_tags: bool = False):
        """The reference CLI's loop over a buffer of raw bytes -> (output bytes, n_lines)."""
        cap = (3 + (64 if predict_tags else 0)) * len(data) + data.count(b"\n") + 16
        buf = C.create_string_buffer(cap)
        nl = C.c_uint64(0)
        mask = sum(1 << ("DRHTKOG".index(ch) + 1) for ch in set(wsconst))
        fn = lib().ora_tokenize_lines_tags if predict_tags else lib().ora_tokenize_lines
        n = fn(self._p, data, len(data), int(no_norm), mask, buf, cap, C.byref(nl))
        if n < 0:
            raise _err(-n)
        return buf.raw[:n], int(nl.value)

    def predict_tags(self, text: str):
        b = text.encode("utf-8")
        cap = max(len(b), 1)
        tt = np.zeros(cap, np.int32)
        ti = np.zeros(cap * max(self.n_tags, 1), np.int32)
        n = lib().ora_predict_tags(self._p, b, len(b), tt.ctypes.data, ti.ctypes.data)
        if n < 0:
            raise _err(-n)
        return tt[:n].copy(), ti[: n * self.n_tags].reshape(n, self.n_tags).copy()

    def add_tag_scores(self, which: int, text: str, token_id: int, pos: int, init):
        b = text.encode("utf-8")
        sc = np.array(init, np.int32)
        rc = lib().ora_add_tag_scores(self._p, which, b, len(b), token_id, pos, sc.ctypes.data, len(sc))
        if rc:
            raise _err(rc)
        return sc

    def dump_patterns(self, which: int):
        cap = 1 << 20
        buf = C.create_string_buffer(cap)
        n = lib().ora_dump_patterns(self._p, which, buf, cap)
        assert n >= 0
        out = []
        for line in buf.raw[:n].split(b"\n"):
            if not line:
                continue
            pat, off, ws = line.split(b"\t")
            if off == b"none":
                out.append((pat, None, None))
            else:
                out.append((pat, int(off), [int(x) for x in ws.split(b",")] if ws else []))
        return out

    def predict_batch(self, text: np.ndarray, offsets: np.ndarray, nthreads: int = 1, want_scores=True):
        """text: uint8 array; offsets: uint64 [n+1]. Returns (scores, boundaries, bound_offsets, status)."""
        L = lib()
        n = len(offsets) - 1
        text = np.ascontiguousarray(text, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        nchars = np.zeros(n, np.uint64)
        L.ora_count_chars(text.ctypes.data, offsets.ctypes.data, n, nchars.ctypes.data)
        nb = np.where(nchars > 0, nchars - 1, 0).astype(np.uint64)
        boff = np.zeros(n + 1, np.uint64)
        np.cumsum(nb, out=boff[1:])
        total = int(boff[-1])
        scores = np.zeros(total, np.int32) if want_scores else None
        bounds = np.zeros(total, np.uint8)
        status = np.zeros(n, np.int32)
        rc = L.ora_predict_batch(self._p, text.ctypes.data, offsets.ctypes.data, n, boff.ctypes.data,
                                 scores.ctypes.data if want_scores else None, bounds.ctypes.data,
                                 status.ctypes.data, nthreads)
        if rc:
            raise _err(rc)
        return scores, bounds, boff, status

    def predict_batch_states(self, text: np.ndarray, offsets: np.ndarray, nthreads: int = 1):
        """Pattern-id states of every character of the batch: (char_states, type_states, char_offsets)."""
        L = lib()
        n = len(offsets) - 1
        text = np.ascontiguousarray(text, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        nchars = np.zeros(n, np.uint64)
        L.ora_count_chars(text.ctypes.data, offsets.ctypes.data, n, nchars.ctypes.data)
        coff = np.zeros(n + 1, np.uint64)
        np.cumsum(nchars, out=coff[1:])
        cs = np.full(int(coff[-1]), 0xFFFFFFFF, np.uint32)
        ts = np.full(int(coff[-1]), 0xFFFFFFFF, np.uint32)
        L.ora_predict_batch_states.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_int]
        rc = L.ora_predict_batch_states(self._p, text.ctypes.data, offsets.ctypes.data, n, coff.ctypes.data,
                                        cs.ctypes.data, ts.ctypes.data, nthreads)
        if rc:
            raise _err(rc)
        return cs, ts, coff

    def time_batch(self, text: np.ndarray, offsets: np.ndarray, nthreads: int = 1):
        """Run predict over the batch discarding outputs placement cost (still computed); returns seconds."""
        import time
        L = lib()
        n = len(offsets) - 1
        status = np.zeros(n, np.int32)
        t0 = time.perf_counter()
        rc = L.ora_predict_batch(self._p, text.ctypes.data, offsets.ctypes.data, n, None, None, None,
                                 status.ctypes.data, nthreads)
        t1 = time.perf_counter()
        if rc:
            raise _err(rc)
        return t1 - t0


def _bench_batch(self, text: np.ndarray, offsets: np.ndarray, nthreads: int = 1, reps: int = 3):
    """Timing loop of the CPU baseline (ora_bench_batch): one pinned thread pool for all `reps` repetitions, dynamic
    sentence blocks; returns the list of per-repetition seconds."""
    L = lib()
    n = len(offsets) - 1
    secs = (C.c_double * reps)()
    L.ora_bench_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
    rc = L.ora_bench_batch(self._p, text.ctypes.data, offsets.ctypes.data, n, nthreads, reps, secs)
    if rc:
        raise _err(rc)
    return list(secs)


OraclePredictor.bench_batch = _bench_batch


def char_types(text: str) -> np.ndarray:
    b = text.encode("utf-8")
    out = np.zeros(max(len(b), 1), np.uint8)
    n = lib().ora_char_types(b, len(b), out.ctypes.data)
    if n < 0:
        raise _err(-n)
    return out[:n].copy()


def kytea_to_model_bytes(data: bytes) -> bytes:
    """KyteaModel::read + Model::try_from + Model::to_vec through the oracle's restatement (kytea_model.rs, model.rs)."""
    L = lib()
    m = C.c_void_p()
    rc = L.ora_model_from_kytea(data, len(data), C.byref(m))
    if rc:
        raise _err(rc)
    try:
        cap = 1 << 16
        while True:
            buf = C.create_string_buffer(cap)
            n = L.ora_model_to_vec(m, buf, cap)
            if n >= 0:
                return buf.raw[:n]
            cap = -n
    finally:
        L.ora_model_free(m)


def grapheme_lengths(text: str):
    """Lengths (in characters) of the extended grapheme clusters of `text`, by the oracle's rule engine."""
    b = text.encode("utf-8")
    lens = np.zeros(max(len(b), 1), np.uint32)
    n = lib().ora_grapheme_lengths(b, len(b), lens.ctypes.data, len(lens))
    if n < 0:
        raise _err(-n)
    return lens[:n].tolist()
