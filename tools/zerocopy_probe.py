#!/usr/bin/env python
"""Experiment: solve straight out of pinned host memory (UVA zero-copy: the kernel
reads/writes host memory over PCIe) vs the chunked copy-engine ring of sks_host_*."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sks_homography_b200 import api, lib  # noqa: E402

L = lib()
dev = torch.device("cuda:0")
n = 1 << 24
src, tar = api.synth_quads(n, 11, 0, torch.float32, dev)
hs = torch.empty((n, 8), dtype=torch.float32, pin_memory=True); hs.copy_(src)
ht = torch.empty((n, 8), dtype=torch.float32, pin_memory=True); ht.copy_(tar)
hH = torch.empty((n, 9), dtype=torch.float32, pin_memory=True)
ref = api.solve("aca", src, tar).cpu()
torch.cuda.synchronize()


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


t = timed(lambda: api.solve("aca", hs, ht, result=hH))
print(f"copy-engine ring : {n / t / 1e9:.3f} GH/s  H2D {n * 64 / t / 1e9:.1f} GB/s  equal={torch.equal(hH, ref)}")
st = torch.cuda.current_stream().cuda_stream
for variant, stages, ctas, name in ((1, 4, 0, "direct"), (2, 4, 0, "ring s4"), (2, 8, 0, "ring s8"), (2, 2, 0, "ring s2")):
    L.c.sks_cuda_set_variant(variant); L.c.sks_cuda_set_tuning(0, stages, ctas)
    hH.zero_()
    f = lambda: L.check(L.c.sks_cuda_aca_f32(hs.data_ptr(), ht.data_ptr(), hH.data_ptr(), n, 0, 0, 1, None, st), "zc")
    t = timed(f)
    print(f"zero-copy {name:8s}: {n / t / 1e9:.3f} GH/s  H2D {n * 64 / t / 1e9:.1f} GB/s  equal={torch.equal(hH, ref)}")
# mixed: inputs zero-copy (kernel reads host), output to device then one D2H copy
L.c.sks_cuda_set_variant(2); L.c.sks_cuda_set_tuning(0, 4, 0)
dH = torch.empty((n, 9), dtype=torch.float32, device=dev)
def mixed():
    L.check(L.c.sks_cuda_aca_f32(hs.data_ptr(), ht.data_ptr(), dH.data_ptr(), n, 0, 0, 1, None, st), "zc")
    hH.copy_(dH, non_blocking=True)
t = timed(mixed)
print(f"zero-copy in, D2H copy out (serial): {n / t / 1e9:.3f} GH/s")
