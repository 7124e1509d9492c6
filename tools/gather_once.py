#!/usr/bin/env python
"""One fused gather+solve launch at the Table-8 replay size (for ncu)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sks_homography_b200 import api
dev = torch.device("cuda:0")
n = 1 << 24
pool = torch.from_numpy(np.random.default_rng(0).uniform(7, 790, size=(2540, 4))).to(dev)
rand4 = api.curand_mrg32k3a(4 * n, 11, dev).view(4, n)
for _ in range(4):
    H = api.gather_solve("aca", pool, n, rand4=rand4, normalize=False, layout="soa")
torch.cuda.synchronize()
print("ok", float(H[0, 0]))
