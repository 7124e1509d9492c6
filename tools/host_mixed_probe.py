#!/usr/bin/env python
"""Experiment: copy-engine H2D for the inputs + kernel writing H straight into the
pinned host result (posted PCIe writes) vs the all-copy-engine ring of sks_host_*."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sks_homography_b200 import api, lib  # noqa: E402
L = lib()
dev = torch.device("cuda:0")
n = 1 << 25
src, tar = api.synth_quads(n, 11, 0, torch.float32, dev)
hs = torch.empty((n, 8), dtype=torch.float32, pin_memory=True); hs.copy_(src)
ht = torch.empty((n, 8), dtype=torch.float32, pin_memory=True); ht.copy_(tar)
hH = torch.empty((n, 9), dtype=torch.float32, pin_memory=True)
ref = api.solve("aca", src, tar).cpu()
del src, tar
torch.cuda.synchronize()

def best_of(fn, reps=4):
    fn(); torch.cuda.synchronize()
    b = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); b = min(b, time.perf_counter() - t0)
    return b

t = best_of(lambda: api.solve("aca", hs, ht, result=hH))
print(f"copy-engine ring (sks_host_*)      : {n / t / 1e9:.3f} G H/s  equal={torch.equal(hH, ref)}")

for chunk_log2 in (19, 20, 21, 22):
    chunk = 1 << chunk_log2
    ring = 3
    streams = [torch.cuda.Stream() for _ in range(ring)]
    dsrc = [torch.empty((chunk, 8), dtype=torch.float32, device=dev) for _ in range(ring)]
    dtar = [torch.empty((chunk, 8), dtype=torch.float32, device=dev) for _ in range(ring)]
    def mixed():
        for ci in range(n // chunk):
            k = ci % ring
            with torch.cuda.stream(streams[k]):
                dsrc[k].copy_(hs[ci * chunk:(ci + 1) * chunk], non_blocking=True)
                dtar[k].copy_(ht[ci * chunk:(ci + 1) * chunk], non_blocking=True)
                out = hH[ci * chunk:(ci + 1) * chunk]
                L.check(L.c.sks_cuda_aca_f32(dsrc[k].data_ptr(), dtar[k].data_ptr(), out.data_ptr(), chunk, 0, 0, 1,
                                             None, streams[k].cuda_stream), "k")
    hH.zero_()
    t = best_of(mixed)
    print(f"CE H2D + kernel writes host, chunk 2^{chunk_log2}: {n / t / 1e9:.3f} G H/s  equal={torch.equal(hH, ref)}")
