#!/usr/bin/env python
"""Small run of every kernel family for compute-sanitizer (tools/sanitize.sh): the direct
AoS kernel (shared-memory output transpose), the TMA ring kernel, SoA, ACA-rect, the
fused gather->solve kernels, the RANSAC scorer in all four modes, finalize, refit and
the warp-grid kernels, at ragged sizes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sks_homography_b200 import api, lib

L = lib()
dev = torch.device("cuda:0")
for tdt in (torch.float32, torch.float64):
    for n in (1, 255, 4099):
        src, tar = api.synth_quads(n, seed=3, dist=1, dtype=tdt, device=dev)
        for variant in (1, 2):
            L.c.sks_cuda_set_variant(variant)
            for solver in ("aca", "sks", "ge"):
                api.solve(solver, src, tar)
        L.c.sks_cuda_set_variant(0)
        api.solve("aca", src.T.contiguous(), tar.T.contiguous(), layout="soa")
        api.aca_rect(tar, 128.0, 1.0, 15.0, 12.0)
corr = api.synth_corr(3, 1001, seed=5, device=dev)
for mode in (0, 1, 2, 3):
    for hpt in (2, 4):
        L.c.sks_cuda_set_ransac_tuning(hpt, 2, mode)
        keys = api.ransac_keys(corr, 1500, seed=7, thr2=2.25)
L.c.sks_cuda_set_ransac_tuning(2, 8, 3)
big = api.synth_corr(2, 9001, seed=5, device=dev)            # multi-tile path
api.ransac_keys(big, 700, seed=7, thr2=2.25)
H, cnt, mask = api.ransac_finalize(corr, 1500, 7, 2.25, keys, want_mask=True)
api.ransac_refit(corr, mask, H)
torch.cuda.synchronize()
print("sanitize_target: done,", int(L.c.sks_cuda_launch_count()), "launches")
