#!/usr/bin/env python
"""Host-pointer path with PAGEABLE caller buffers (what a std::vector-holding C++ caller
of the reference has) vs pinned ones; staging copies with non-temporal stores vs memcpy,
and the staging chunk size."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sks_homography_b200 import api, lib  # noqa: E402
L = lib()
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
ref = None
for nt in (3, 0, 1, 2, 3):
    L.c.sks_host_set_staging_copy(nt)
    tp = best(lambda: api.solve("aca", s, t, result=H))
    if ref is None:
        ref = H.clone()
    print(f"pageable in/out, staging_copy={nt}: {n / tp / 1e9:.3f} G H/s  same={torch.equal(H.view(torch.int32), ref.view(torch.int32))}", flush=True)
L.c.sks_host_set_staging_copy(3)
for parts in (2, 4, 6, 8, 12, 16):
    L.c.sks_host_set_staging_threads(parts)
    tp = best(lambda: api.solve("aca", s, t, result=H))
    print(f"pageable in/out, NT, {parts:2d} threads per copy: {n / tp / 1e9:.3f} G H/s", flush=True)
L.c.sks_host_set_staging_threads(0)
sp, tq, Hp = s.pin_memory(), t.pin_memory(), H.pin_memory()
tp = best(lambda: api.solve("aca", sp, tq, result=Hp))
print(f"pinned in/out  : {n / tp / 1e9:.3f} G H/s")
for nt in (3, 0):
    L.c.sks_host_set_staging_copy(nt)
    tp = best(lambda: api.solve("aca", sp, tq, result=H))
    print(f"pinned in, pageable out ({'NT' if nt else 'memcpy'}): {n / tp / 1e9:.3f} G H/s")
    tp = best(lambda: api.solve("aca", s, t, result=Hp))
    print(f"pageable in, pinned out ({'NT' if nt else 'memcpy'}): {n / tp / 1e9:.3f} G H/s")
