#!/usr/bin/env python
"""One fused ACA-rect -> warp-grid launch (16384 samples x 128x128 points) for ncu."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sks_homography_b200 import api
dev = torch.device("cuda:0")
n, g = 16384, 128
_, tar = api.synth_quads(n, seed=11, device=dev)
out = torch.empty((n, g, g, 2), dtype=torch.float32, device=dev)
for _ in range(4):
    api.aca_rect_warp_grid(tar, 128.0, 1.0, g, g, M_x=15.0, M_y=12.0, out=out, x0=15.0, y0=12.0, dx=1.0, dy=1.0)
torch.cuda.synchronize()
print("ok", float(out[0, 0, 0, 0]))
