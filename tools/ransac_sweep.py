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
CAP2 = "-DSKS_RANSAC_MIN_CTAS(T)=2"
VARIANTS = {"rs_base": [], "rs_fp_h3": ["-DSKS_RANSAC_HPT_HI=3"], "rs_fp_h3_u1": ["-DSKS_RANSAC_HPT_HI=3", "-DSKS_RANSAC_FP_UNROLL=1"],
            "rs_fp_h3_u3": ["-DSKS_RANSAC_HPT_HI=3", "-DSKS_RANSAC_FP_UNROLL=3"]}
EXTRA = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")]
for e in EXTRA:                      # --variant=tag:-DX=1,-DY=2
    tag, defs = e.split(":", 1)
    VARIANTS[tag] = defs.split(",")

if "--build" in sys.argv:
    from sks_homography_b200 import build as b
    os.makedirs(VDIR, exist_ok=True)
    for f in glob.glob(os.path.join(VDIR, "*.so")):
        os.remove(f)
    for tag, defs in VARIANTS.items():
        out = os.path.join(VDIR, f"libsks_cuda_{tag}.so")
        cmd = [b.nvcc()] + b.NVCC_FLAGS + defs + ["-Xptxas", "-v", "-o", out] + [os.path.join(b.CSRC, f) for f in b.SOURCES]
        r = subprocess.run(cmd, capture_output=True, text=True)
        regs = {}
        cur = None
        for l in r.stderr.splitlines():
            if "Compiling entry function" in l and "k_ransac_aca" in l:
                cur = l.split("k_ransac_acaILi")[1][:14]
            elif "Used" in l and cur:
                regs[cur] = l.split("Used ")[1].split(" ")[0]; cur = None
        print(tag, "ok" if r.returncode == 0 else r.stderr[-400:], {k: v for k, v in regs.items() if "ELi3ELi256" in k or "ELi1ELi256" in k})
    sys.exit(0)

import torch
from sks_homography_b200 import _lib, api

dev = torch.device("cuda:0")
P, n_pts, n_hyp = 256, 4096, 65536
corr = api.synth_corr(P, n_pts, seed=11, device=dev)
ref = api.ransac_keys(corr, n_hyp, 11, 2.25)
st = torch.cuda.current_stream().cuda_stream
CONFIGS = [(3, 2, 0), (3, 4, 0), (3, 4, 1)]
for path in sorted(glob.glob(os.path.join(VDIR, "libsks_cuda_rs_*.so"))):
    L = _lib.SksCuda(path)
    for mode, hpt, thr in CONFIGS:
        if "_fp" in path and mode != 3:
            continue
        for rounds in ((8,) if hpt == 2 else (4,)):
            L.c.sks_cuda_set_ransac_tuning(hpt, rounds, mode | (thr << 2))
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
            print(f"{os.path.basename(path):24s} mode={mode} hpt={hpt} rounds={rounds} threads={(256, 384, 512)[thr]} "
                  f"{ms:8.3f} ms  {tf:6.2f} TFLOP/s  {tf / 74.45:.3f} of peak  same_keys={ok}", flush=True)
