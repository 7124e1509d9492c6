#!/usr/bin/env python
"""One fused ACA-RANSAC launch for ncu: python tools/ransac_once.py [pairs] [mode] [hpt] [threads-code]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sks_homography_b200 import api, lib
L = lib()
P = int(sys.argv[1]) if len(sys.argv) > 1 else 444
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 3
hpt = int(sys.argv[3]) if len(sys.argv) > 3 else 2
thr = int(sys.argv[4]) if len(sys.argv) > 4 else 0
L.c.sks_cuda_set_ransac_tuning(hpt, 8 if hpt == 2 else 4, mode | (thr << 2))
dev = torch.device("cuda:0")
corr = api.synth_corr(P, 4096, seed=11, device=dev)
for _ in range(2):
    keys = api.ransac_keys(corr, 65536, 11, 2.25)
torch.cuda.synchronize()
print("done", int(keys[0].item()) >> 32)
