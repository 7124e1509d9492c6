#!/usr/bin/env python
"""Consumer after the path: ACA-rect -> sampling grid, fused (H stays in registers) vs
solver kernel + grid kernel (H through HBM).  n samples x gh x gw points, fp32."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sks_homography_b200 import api

dev = torch.device("cuda:0")


def time_ms(fn, it=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


for n, g in ((4096, 128), (16384, 128), (65536, 32), (262144, 32), (1 << 20, 8), (1 << 22, 8), (1 << 22, 2)):
    _, tar = api.synth_quads(n, seed=11, device=dev)
    out = torch.empty((n, g, g, 2), dtype=torch.float32, device=dev)
    H = torch.empty((n, 9), dtype=torch.float32, device=dev)
    spec = dict(x0=15.0, y0=12.0, dx=128.0 / (g - 1), dy=128.0 / (g - 1))

    def two():
        api.aca_rect(tar, 128.0, 1.0, 15.0, 12.0, result=H, normalize=False)
        api.warp_grid(H, g, g, out=out, **spec)
    fused = lambda: api.aca_rect_warp_grid(tar, 128.0, 1.0, g, g, M_x=15.0, M_y=12.0, out=out, **spec)
    t2, t1 = time_ms(two), time_ms(fused)
    gb = (n * g * g * 8 + n * 32) / 1e9
    print(f"n={n:8d} grid {g:3d}x{g:<3d}: solver + grid kernels {1e3 * t2:9.1f} us | fused {1e3 * t1:9.1f} us "
          f"({t2 / t1:.2f}x)  fused = {gb / (t1 * 1e-3):7.0f} GB/s of grid writes + corner reads", flush=True)
