#!/usr/bin/env python
"""Host-pointer path with PAGEABLE caller buffers (what a std::vector-holding C++ caller
of the reference has) vs pinned ones."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sks_homography_b200 import api  # noqa: E402
n = 1 << 25
src, tar = api.synth_quads(n, 11, 0, torch.float32, torch.device("cuda:0"))
s, t = src.cpu(), tar.cpu()
del src, tar
H = torch.empty((n, 9), dtype=torch.float32)
def best(fn, reps=3):
    fn(); b = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); b = min(b, time.perf_counter() - t0)
    return b
tp = best(lambda: api.solve("aca", s, t, result=H))
print(f"pageable in/out: {n / tp / 1e9:.3f} G H/s")
sp, tq, Hp = s.pin_memory(), t.pin_memory(), H.pin_memory()
tp = best(lambda: api.solve("aca", sp, tq, result=Hp))
print(f"pinned in/out  : {n / tp / 1e9:.3f} G H/s")
tp = best(lambda: api.solve("aca", sp, tq, result=H))
print(f"pinned in, pageable out: {n / tp / 1e9:.3f} G H/s")
