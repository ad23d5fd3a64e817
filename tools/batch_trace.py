"""Manual tuning aid (run on the GPU box): vpt_predict_batch end-to-end time + pipeline trace (VPT_TRACE)."""
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
    mb = bench.get_model(300_000, 2_000_000, 2)
    pred = vb.Predictor(vb.Model.read(mb))
    text, offs, _ = bench.get_text(n, 0, False)
    nb = int(offs[-1])
    n_bound = nb  # upper bound
    h_text = torch.from_numpy(text).pin_memory()
    h_off = torch.from_numpy(offs.astype(np.int64)).pin_memory()
    h_scores = torch.empty(n_bound, dtype=torch.int32).pin_memory()
    h_bounds = torch.empty(n_bound, dtype=torch.uint8).pin_memory()
    h_boff = torch.empty(n + 1, dtype=torch.int64).pin_memory()
    h_status = torch.empty(n, dtype=torch.int32).pin_memory()
    a, b = C.c_uint64(), C.c_uint64()
    L = vb.lib()

    def step(scores=True):
        rc = L.vpt_predict_batch(pred._h, h_text.data_ptr(), h_off.data_ptr(), n, h_scores.data_ptr() if scores else None,
                                 h_bounds.data_ptr(), n_bound, h_boff.data_ptr(), h_status.data_ptr(), None, None, 0, None,
                                 C.byref(a), C.byref(b))
        assert rc == 0, L.vpt_last_error()

    for scores in (True, False):
        for _ in range(3):
            step(scores)
        t0 = time.perf_counter()
        for _ in range(5):
            step(scores)
        dt = (time.perf_counter() - t0) / 5
        print(f"predict_batch scores={scores}: {dt * 1e3:.3f} ms  {nb / dt / 1e9:.2f} GB/s", flush=True)
        os.environ["VPT_TRACE"] = "1"
        step(scores)
        os.environ["VPT_TRACE"] = "0"


if __name__ == "__main__":
    main()
