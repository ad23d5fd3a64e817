"""Manual tuning aid (run on the GPU box): vpt_tokenize_lines throughput against VPT_CHUNK_BYTES."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import vaporetto_b200 as vb  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    chunks = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2, 4, 8, 16, 32]
    mb = bench.get_model(300_000, 2_000_000, 2)
    pred = vb.Predictor(vb.Model.read(mb))
    text, offs, _ = bench.get_text(n, 0, False)
    nbytes = int(offs[-1])
    starts = offs[:-1].astype(np.int64) + np.arange(n, dtype=np.int64)
    lines = np.full(nbytes + n, 10, np.uint8)
    keep = np.ones(nbytes + n, bool)
    keep[starts[1:] - 1] = False
    keep[-1] = False
    lines[keep] = text[:nbytes]
    h_in = torch.from_numpy(lines).pin_memory()
    h_out = torch.empty(3 * (nbytes + n), dtype=torch.uint8).pin_memory()
    ln, nl = C.c_uint64(), C.c_uint64()
    L = vb.lib()
    for mbs in chunks:
        os.environ["VPT_CHUNK_BYTES"] = str(mbs << 20)
        for _ in range(2):
            assert L.vpt_tokenize_lines(pred._h, h_in.data_ptr(), nbytes + n, 1, 0, h_out.data_ptr(), h_out.numel(), C.byref(ln), C.byref(nl)) == 0
        t0 = time.perf_counter()
        k = 5
        for _ in range(k):
            L.vpt_tokenize_lines(pred._h, h_in.data_ptr(), nbytes + n, 1, 0, h_out.data_ptr(), h_out.numel(), C.byref(ln), C.byref(nl))
        dt = (time.perf_counter() - t0) / k
        if os.environ.get("TRACE_ONE"):
            os.environ["VPT_TRACE"] = "1"
            L.vpt_tokenize_lines(pred._h, h_in.data_ptr(), nbytes + n, 1, 0, h_out.data_ptr(), h_out.numel(), C.byref(ln), C.byref(nl))
            os.environ["VPT_TRACE"] = "0"
        print(f"chunk {mbs:3d} MiB: {dt * 1e3:7.3f} ms  {nbytes / dt / 1e9:6.2f} GB/s  out {ln.value} bytes, {nl.value} lines", flush=True)


if __name__ == "__main__":
    main()
