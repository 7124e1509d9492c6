#!/usr/bin/env python
"""Experiment: cache-hint / CTA-size variants of the direct AoS kernel.
`--build` (CPU box) compiles libsks_cuda_<tag>.so with different -D knobs into
gpurun_out-free tools/_variants/; without it (GPU box) times each on ACA f32 2^26."""
import ctypes as C
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "tools", "_variants")
VARIANTS = {"A_ld0_st4": ["-DSKS_WLD_HINT=0", "-DSKS_WST_HINT=0"],
            "B_ld1_st4": ["-DSKS_WLD_HINT=1", "-DSKS_WST_HINT=0"],
            "C_ld2_st4": ["-DSKS_WLD_HINT=2", "-DSKS_WST_HINT=0"],
            "D_ld0_st8plain": ["-DSKS_WLD_HINT=0", "-DSKS_WST_HINT=1"],
            "E_ld0_st8na": ["-DSKS_WLD_HINT=0", "-DSKS_WST_HINT=2"],
            "F_ld2_st8ef": ["-DSKS_WLD_HINT=2", "-DSKS_WST_HINT=3"],
            "G_ld1_st8na": ["-DSKS_WLD_HINT=1", "-DSKS_WST_HINT=2"],
            "H_ld2_st4_t128": ["-DSKS_WLD_HINT=2", "-DSKS_WST_HINT=0", "-DSKS_DIRECT_TILE_F32=128"],
            "I_ld2_st4_t512": ["-DSKS_WLD_HINT=2", "-DSKS_WST_HINT=0", "-DSKS_DIRECT_TILE_F32=512"]}

if "--build" in sys.argv:
    from sks_homography_b200 import build as b
    os.makedirs(VDIR, exist_ok=True)
    for tag, defs in VARIANTS.items():
        out = os.path.join(VDIR, f"libsks_cuda_{tag}.so")
        cmd = [b.nvcc()] + b.NVCC_FLAGS + defs + ["-o", out] + [os.path.join(b.CSRC, f) for f in b.SOURCES]
        r = subprocess.run(cmd, capture_output=True, text=True)
        print(tag, "ok" if r.returncode == 0 else r.stderr[-400:])
    sys.exit(0)

import torch
from sks_homography_b200 import _lib, api

dev = torch.device("cuda:0")
n = 1 << 26
src, tar = api.synth_quads(n, 11, 0, torch.float32, dev)
H = torch.empty((n, 9), dtype=torch.float32, device=dev)
ref = None
st = torch.cuda.current_stream().cuda_stream
for path in sorted(glob.glob(os.path.join(VDIR, "*.so"))):
    L = _lib.SksCuda(path)
    for solver, narrow in (("aca", 0), ("aca", 2), ("sks", 0), ("sks", 2), ("rect", 0), ("rect", 2)):
        L.c.sks_cuda_set_tuning(narrow, 4, 0)
        if solver == "rect":
            run = lambda: L.check(L.c.sks_cuda_aca_rect_f32(tar.data_ptr(), None, 15.0, 12.0, 128.0, 1.0,
                                                            H.data_ptr(), n, 0, 0, 1, None, st), "run")
        else:
            fn = getattr(L.c, f"sks_cuda_{solver}_f32")
            run = lambda: L.check(fn(src.data_ptr(), tar.data_ptr(), H.data_ptr(), n, 0, 0, 1, None, st), "run")
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
        ev[0].record()
        for i in range(20):
            run(); ev[i + 1].record()
        torch.cuda.synchronize()
        ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(20))
        print(f"{os.path.basename(path):36s} {solver} {'v4 ' if narrow else 'v8 '} median {ts[10]:.4f} ms  min {ts[0]:.4f}  {n * (68 if solver == 'rect' else 100) / ts[10] / 1e6:.1f} GB/s", flush=True)
