#!/usr/bin/env python
"""Experiment: does the power-of-two plane stride of the SoA layout (ld = n = 2^26)
alias DRAM channels?  Same solve with ld = n + pad."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sks_homography_b200 import api, lib  # noqa: E402
L = lib()
dev = torch.device("cuda:0")
n = 1 << 26
st = torch.cuda.current_stream().cuda_stream
for pad in (0, 256, 4096, 65536 + 256, (1 << 20) + 4096):
    ld = n + pad
    src = torch.empty((8, ld), dtype=torch.float32, device=dev)
    tar = torch.empty((8, ld), dtype=torch.float32, device=dev)
    H = torch.empty((9, ld), dtype=torch.float32, device=dev)
    L.check(L.c.sks_cuda_synth_quads_f32(src.data_ptr(), tar.data_ptr(), 0, n, 11, 0, 1, ld, st), "synth")
    run = lambda: L.check(L.c.sks_cuda_aca_f32(src.data_ptr(), tar.data_ptr(), H.data_ptr(), n, 1, ld, 1, None, st), "soa")
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    ev[0].record()
    for i in range(10):
        run(); ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(10))
    print(f"SoA ACA f32 n=2^26 ld=n+{pad:8d}: {ts[5]:.4f} ms  {n * 100 / ts[5] / 1e6:.1f} GB/s", flush=True)
    del src, tar, H
