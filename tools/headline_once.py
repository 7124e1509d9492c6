#!/usr/bin/env python
"""The headline launch (ACA fp32, 2^26 quadruples, AoS, h33-normalised) a few times, for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sks_homography_b200 import api
dev = torch.device("cuda:0")
n = 1 << 26
src, tar = api.synth_quads(n, seed=11, dist=0, dtype=torch.float32, device=dev)
H = torch.empty((n, 9), dtype=torch.float32, device=dev)
for _ in range(4):
    api.solve("aca", src, tar, result=H)
torch.cuda.synchronize()
print("done", float(H[0, 8]))
