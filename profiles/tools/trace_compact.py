#!/usr/bin/env python
"""Pipeline timeline of vpt_predict_batch_compact / vpt_predict_batch (VPT_TRACE=1) on the config-2 workload.
usage (GPU box): python profiles/tools/trace_compact.py [n_sentences] [tags] 2> trace.txt"""
import os
import sys

os.environ.setdefault("VPT_TRACE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C
import time

import numpy as np
import torch

import vaporetto_b200 as vb
from vpt_testlib import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
tags = len(sys.argv) > 2 and sys.argv[2] == "tags"   # config 3: 20 000 tag models, tag prediction on the device
mb = synth.gen_model_bccwj_shaped(n_patterns=300_000, sample_sentences=2_000_000, tag_models=20_000 if tags else 0)
pred = vb.Predictor(vb.Model.read(mb), predict_tags=tags)
text, offs, _ = synth.gen_text(n, 40)
L = vb.lib()
h_text = torch.from_numpy(np.asarray(text)).pin_memory()
h_off = torch.from_numpy(offs.astype(np.int64)).pin_memory()
nb = int(offs[-1])
h_bits = torch.zeros(nb // 32 + 2, dtype=torch.int32).pin_memory()
h_nch = torch.empty(n, dtype=torch.int32).pin_memory()
h_st = torch.empty(n, dtype=torch.uint8).pin_memory()
h_ntok = torch.empty(n, dtype=torch.int32).pin_memory()
h_tid = torch.empty(nb if tags else 1, dtype=torch.int32).pin_memory()
h_cand = torch.empty(nb * max(pred.n_tags, 1) if tags else 1, dtype=torch.uint8).pin_memory()
a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
for it in range(3):
    t0 = time.perf_counter()
    rc = L.vpt_predict_batch_compact(pred._h, h_text.data_ptr(), h_off.data_ptr(), n, h_bits.data_ptr(), h_bits.numel(),
                                     h_nch.data_ptr(), h_st.data_ptr(), h_ntok.data_ptr(), h_tid.data_ptr() if tags else None,
                                     h_cand.data_ptr() if tags else None, nb if tags else 0, C.byref(a), C.byref(b), C.byref(c))
    assert rc == 0, L.vpt_last_error()
    print("compact call %d: %.3f ms" % (it, 1e3 * (time.perf_counter() - t0)), file=sys.stderr)
