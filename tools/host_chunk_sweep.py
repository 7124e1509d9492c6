#!/usr/bin/env python
"""e2e host-pointer path vs pipeline chunk size (pinned buffers, one GPU)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sks_homography_b200 import api, lib  # noqa: E402
L = lib()
for log2n in (25, 26):
    n = 1 << log2n
    src, tar = api.synth_quads(n, 11, 0, torch.float32, torch.device("cuda:0"))
    hs = torch.empty((n, 8), dtype=torch.float32, pin_memory=True); hs.copy_(src)
    ht = torch.empty((n, 8), dtype=torch.float32, pin_memory=True); ht.copy_(tar)
    hH = torch.empty((n, 9), dtype=torch.float32, pin_memory=True)
    del src, tar
    for mb in (1, 2, 4, 8, 16, 32, 64, 128):
        L.c.sks_host_set_chunk_bytes(mb << 20)
        api.solve("aca", hs, ht, result=hH)
        best = 1e9
        for _ in range(4):
            t0 = time.perf_counter(); api.solve("aca", hs, ht, result=hH); best = min(best, time.perf_counter() - t0)
        print(f"n=2^{log2n} chunk {mb:4d} MiB/array: {n / best / 1e9:.3f} G H/s  H2D {n * 64 / best / 1e9:.1f} GB/s", flush=True)
    del hs, ht, hH
