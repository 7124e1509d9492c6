#!/usr/bin/env python
"""e2e throughput of the host-pointer path when ONE process shards the batch over
g GPUs itself (sks_host_set_device_count), pinned buffers, 2^26 quadruples."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sks_homography_b200 import api, lib  # noqa: E402

L = lib()
n = 1 << 26
src, tar = api.synth_quads(n, 11, 0, torch.float32, torch.device("cuda:0"))
hs = torch.empty((n, 8), dtype=torch.float32, pin_memory=True); hs.copy_(src)
ht = torch.empty((n, 8), dtype=torch.float32, pin_memory=True); ht.copy_(tar)
hH = torch.empty((n, 9), dtype=torch.float32, pin_memory=True)
ref = api.solve("aca", src, tar).cpu()
del src, tar
for g in [x for x in (1, 2, 4, 8) if x <= torch.cuda.device_count()]:
    L.c.sks_host_set_device_count(g)
    api.solve("aca", hs, ht, result=hH)
    t0 = time.perf_counter()
    for _ in range(3):
        api.solve("aca", hs, ht, result=hH)
    el = (time.perf_counter() - t0) / 3
    print(f"in-process host driver, {g} GPU(s): {n / el / 1e9:.3f} G H/s  ({n * 64 / el / 1e9:.1f} GB/s H2D aggregate)  "
          f"equal={torch.equal(hH, ref)}", flush=True)
