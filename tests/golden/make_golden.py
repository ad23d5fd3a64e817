#!/usr/bin/env python
"""Copy / derive the reference's binary fixtures for the predict path into tests/golden/.

Run in the build container (needs /root/reference).  The GPU box has no
/root/reference, so the outputs are committed.

  resources/model.bin                    -> model.bin             (raw Vaporetto model, 394 B)
  resources/docs.tok                     -> docs.tok
  resources/kytea-model.bin              -> kytea-model.bin       (KyTea binary model, converter fixture)
  vaporetto_tantivy/test_model/model.zst -> tantivy_model.bin     (zstd-decompressed via libzstd.so.1)
  vaporetto_tantivy/test_model/test_corpus.tok -> tantivy_test_corpus.tok
"""
import ctypes
import os
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def unzstd(data: bytes) -> bytes:
    z = ctypes.CDLL("libzstd.so.1")
    z.ZSTD_getFrameContentSize.restype = ctypes.c_ulonglong
    z.ZSTD_getFrameContentSize.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
    z.ZSTD_decompress.restype = ctypes.c_size_t
    z.ZSTD_decompress.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
    size = z.ZSTD_getFrameContentSize(data, len(data))
    if size >= (1 << 62):
        size = 1 << 20
    buf = ctypes.create_string_buffer(size)
    n = z.ZSTD_decompress(buf, size, data, len(data))
    assert n <= size
    return buf.raw[:n]


def main():
    shutil.copyfile(f"{REF}/resources/model.bin", f"{HERE}/model.bin")
    shutil.copyfile(f"{REF}/resources/docs.tok", f"{HERE}/docs.tok")
    shutil.copyfile(f"{REF}/resources/kytea-model.bin", f"{HERE}/kytea-model.bin")
    shutil.copyfile(f"{REF}/vaporetto_tantivy/test_model/test_corpus.tok", f"{HERE}/tantivy_test_corpus.tok")
    with open(f"{REF}/vaporetto_tantivy/test_model/model.zst", "rb") as f:
        raw = unzstd(f.read())
    with open(f"{HERE}/tantivy_model.bin", "wb") as f:
        f.write(raw)
    for fn in ["model.bin", "docs.tok", "kytea-model.bin", "tantivy_model.bin", "tantivy_test_corpus.tok"]:
        print(fn, os.path.getsize(f"{HERE}/{fn}"))
        os.chmod(f"{HERE}/{fn}", 0o644)


if __name__ == "__main__":
    main()
