"""Manual experiment (run on the GPU box): does PCIe DMA traffic slow the scoring kernels down?
Times vpt_predict_batch_dev (device-resident, CUDA events) alone and while pinned H2D / D2H copies run on other streams."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import vaporetto_b200 as vb  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    dev = torch.device("cuda:0")
    mb = bench.get_model(300_000, 2_000_000, 2)
    pred = vb.Predictor(vb.Model.read(mb))
    text, offs, _ = bench.get_text(n, 0, False)
    nbytes = int(offs[-1])
    L = vb.lib()
    d_text = torch.zeros(nbytes + 64, dtype=torch.uint8, device=dev)
    d_text[:nbytes] = torch.from_numpy(text[:nbytes]).to(dev)
    d_off = torch.from_numpy(offs.astype(np.int64)).to(dev)
    ws_bytes = L.vpt_workspace_size(n)
    d_ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    d_scores = torch.empty(nbytes, dtype=torch.int32, device=dev)
    d_bounds = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_boff = torch.empty(n + 1, dtype=torch.int64, device=dev)
    d_status = torch.empty(n, dtype=torch.int32, device=dev)
    st = torch.cuda.Stream()

    def step():
        rc = L.vpt_predict_batch_dev(pred._h, d_text.data_ptr(), d_off.data_ptr(), n, d_ws.data_ptr(), ws_bytes,
                                     d_scores.data_ptr(), d_bounds.data_ptr(), d_boff.data_ptr(), d_status.data_ptr(),
                                     None, None, None, st.cuda_stream)
        assert rc == 0, L.vpt_last_error()

    def timed(k=20):
        for _ in range(3):
            step()
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(k):
            step()
        e1.record(st)
        st.synchronize()
        return e0.elapsed_time(e1) / k

    m = 256 << 20
    h1 = torch.empty(m, dtype=torch.uint8).pin_memory()
    h2 = torch.empty(m, dtype=torch.uint8).pin_memory()
    g1 = torch.empty(m, dtype=torch.uint8, device=dev)
    g2 = torch.empty(m, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def dma(h2d, d2h, reps=12):
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    g1.copy_(h1, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h2.copy_(g2, non_blocking=True)

    print(f"alone:            {timed():.4f} ms/step", flush=True)
    for name, a, b in (("with H2D", True, False), ("with D2H", False, True), ("with H2D + D2H", True, True)):
        dma(a, b)
        t = timed()
        busy = not (s1.query() and s2.query())
        torch.cuda.synchronize()
        print(f"{name:17s} {t:.4f} ms/step   (copies still running at the end: {busy})", flush=True)


if __name__ == "__main__":
    main()
