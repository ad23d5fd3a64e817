#!/usr/bin/env python
"""H2D copy rate from pinned memory by copy size (one stream, back to back), alone and while a kernel streams HBM on
another stream.  usage (GPU box): python profiles/tools/h2d_sizes.py"""
import time
import torch

dev = torch.device("cuda:0")
big = 256 << 20
h = torch.empty(big, dtype=torch.uint8).pin_memory()
d = torch.empty(big, dtype=torch.uint8, device=dev)
x = torch.empty(1 << 28, dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(size, busy):
    n = max(1, big // size)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(s1):
        for i in range(n):
            d[i * size:(i + 1) * size].copy_(h[i * size:(i + 1) * size], non_blocking=True)
    if busy:
        with torch.cuda.stream(s2):
            for _ in range(6):
                x.add_(1.0)
    s1.synchronize()
    dt = time.perf_counter() - t0
    torch.cuda.synchronize()
    return n * size / dt / 1e9


for size_mb in (1, 4, 8, 16, 23, 32, 64, 256):
    size = size_mb << 20
    run(size, False)
    print("H2D %4d MB copies: %.1f GB/s alone, %.1f GB/s with an HBM-streaming kernel" % (size_mb, run(size, False), run(size, True)))
