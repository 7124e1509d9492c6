#!/usr/bin/env python
"""Experiment: inner-loop shapes of the packed RANSAC scorer.  `--build` (CPU box)
compiles variants into tools/_variants/; on the GPU box each is timed on
256 pairs x 4096 matches x 65536 hypotheses and checked against the default."""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "tools", "_variants")
VARIANTS = {}
for u in (2, 4, 8):
    VARIANTS[f"rs_u{u}"] = [f"-DSKS_RANSAC_UNROLL={u}"]

if "--build" in sys.argv:
    from sks_homography_b200 import build as b
    os.makedirs(VDIR, exist_ok=True)
    for tag, defs in VARIANTS.items():
        out = os.path.join(VDIR, f"libsks_cuda_{tag}.so")
        cmd = [b.nvcc()] + b.NVCC_FLAGS + defs + ["-Xptxas", "-v", "-o", out] + [os.path.join(b.CSRC, f) for f in b.SOURCES]
        r = subprocess.run(cmd, capture_output=True, text=True)
        regs = [l for l in r.stderr.splitlines() if "registers" in l]
        print(tag, "ok" if r.returncode == 0 else r.stderr[-400:])
    sys.exit(0)

import torch
from sks_homography_b200 import _lib, api

dev = torch.device("cuda:0")
P, n_pts, n_hyp = 256, 4096, 65536
corr = api.synth_corr(P, n_pts, seed=11, device=dev)
ref = api.ransac_keys(corr, n_hyp, 11, 2.25)
st = torch.cuda.current_stream().cuda_stream
for path in sorted(glob.glob(os.path.join(VDIR, "libsks_cuda_rs_*.so"))):
    L = _lib.SksCuda(path)
    for mode, hpt, thr in [(1, 2, 0), (2, 2, 0), (2, 4, 0), (2, 2, 1), (2, 4, 1), (2, 2, 2), (2, 4, 2)]:
        L.c.sks_cuda_set_ransac_tuning(hpt, 8 if hpt == 2 else 4, mode | (thr << 2))
        keys = torch.zeros(P, dtype=torch.int64, device=dev)
        run = lambda: L.check(L.c.sks_cuda_ransac_aca_f32(corr.data_ptr(), P, n_pts, None, n_hyp, 0, n_hyp, 11,
                                                          2.25, keys.data_ptr(), st), "ransac")
        run(); torch.cuda.synchronize()
        ok = bool(torch.equal(keys, ref))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        tf = P * n_hyp * (103 + 21.0 * n_pts) / ms / 1e9
        print(f"{os.path.basename(path):24s} mode={mode} hpt={hpt} threads={(256, 384, 512)[thr]} {ms:8.3f} ms  "
              f"{tf:6.2f} TFLOP/s  {tf / 74.45:.3f} of peak  same_keys={ok}", flush=True)
