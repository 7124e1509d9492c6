#!/usr/bin/env python
"""TensorACA_rect on the reference's [bs,3,4] tensors, a few launches, for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sks_homography_b200 import api
dev = torch.device("cuda:0")
bs = 1 << 24
torch.manual_seed(11)
src = torch.randint(10, 30, (bs, 2), device=dev).float().unsqueeze(1).repeat(1, 4, 1)
src[:, 1, 0] += 128; src[:, 2, 1] += 128; src[:, 3, 0] += 128; src[:, 3, 1] += 128
tar = src + torch.randint(0, 32, (bs, 4, 2), device=dev).float()
ones = torch.ones((bs, 1, 4), device=dev)
src34 = torch.cat((src.transpose(1, 2), ones), dim=1).contiguous()
tar34 = torch.cat((tar.transpose(1, 2), ones), dim=1).contiguous()
scale = src34[0, 0, 1:2] - src34[0, 0, 0:1]
div = scale / (src34[0, 1, 2:3] - src34[0, 1, 0:1])
for _ in range(4):
    H = api.TensorACA_rect(bs, src34, tar34, scale, div)
torch.cuda.synchronize()
print("done", float(H[0, 2, 2]))
