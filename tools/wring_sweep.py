#!/usr/bin/env python
"""Experiment: shapes of the warp-private TMA ring (variant 3) against the direct kernel on the
headline workload (ACA fp32, 2^26 quadruples) and on SKS fp64 2^25.  `--build` (CPU box) compiles
the shapes into tools/_variants/; on the GPU box each is timed and bit-compared with the direct kernel."""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "tools", "_variants")
SHAPES = [(64, 4, 8), (64, 3, 8), (32, 4, 16), (32, 6, 8), (64, 2, 16), (128, 2, 8), (64, 6, 4)]   # QPW_F32, STAGES, WARPS

if "--build" in sys.argv:
    from sks_homography_b200 import build as b
    os.makedirs(VDIR, exist_ok=True)
    for f in glob.glob(os.path.join(VDIR, "libsks_cuda_wr_*.so")):
        os.remove(f)
    for q, st, w in SHAPES:
        out = os.path.join(VDIR, f"libsks_cuda_wr_q{q}_s{st}_w{w}.so")
        cmd = [b.nvcc()] + b.NVCC_FLAGS + [f"-DSKS_WRING_QPW_F32={q}", f"-DSKS_WRING_STAGES={st}", f"-DSKS_WRING_WARPS={w}",
                                           "-o", out] + [os.path.join(b.CSRC, f) for f in b.SOURCES]
        r = subprocess.run(cmd, capture_output=True, text=True)
        print(os.path.basename(out), "ok" if r.returncode == 0 else r.stderr[-300:])
    sys.exit(0)

import torch
from sks_homography_b200 import _lib, api

dev = torch.device("cuda:0")


def time_ms(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return ts[len(ts) // 2]


for solver, tdt, log2n, bph in (("aca", torch.float32, 26, 100), ("sks", torch.float64, 25, 200), ("rect", torch.float32, 26, 68)):
    n = 1 << log2n
    src, tar = api.synth_quads(n, 11, 0 if tdt == torch.float32 else 1, tdt, dev)
    H = torch.empty((n, 9), dtype=tdt, device=dev)
    Href = torch.empty_like(H)
    base = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream

    def run(L, out):
        if solver == "rect":
            fn = L.c.sks_cuda_aca_rect_f32
            L.check(fn(tar.data_ptr(), None, 15.0, 12.0, 128.0, 1.0, out.data_ptr(), n, 0, 0, 1, None, st), "rect")
        else:
            fn = getattr(L.c, f"sks_cuda_{solver}_{'f32' if tdt == torch.float32 else 'f64'}")
            L.check(fn(src.data_ptr(), tar.data_ptr(), out.data_ptr(), n, 0, 0, 1, None, st), solver)

    base.c.sks_cuda_set_variant(1)
    t = time_ms(lambda: run(base, Href))
    print(f"{solver:5s} {str(tdt)[6:]:8s} direct                      {t:7.4f} ms  {n * bph / t / 1e6:7.1f} GB/s", flush=True)
    for path in sorted(glob.glob(os.path.join(VDIR, "libsks_cuda_wr_*.so"))):
        L = _lib.SksCuda(path)
        for ctas in (0, 1):
            L.c.sks_cuda_set_variant(3)
            L.c.sks_cuda_set_tuning(0, 4, ctas)
            H.zero_()
            try:
                t = time_ms(lambda: run(L, H))
            except Exception as e:
                print(f"{solver:5s} {os.path.basename(path)[14:-3]:18s} ctas={ctas}  {type(e).__name__}: {str(e)[:80]}")
                continue
            same = torch.equal(H.view(torch.int64 if tdt == torch.float64 else torch.int32),
                               Href.view(torch.int64 if tdt == torch.float64 else torch.int32))
            print(f"{solver:5s} {str(tdt)[6:]:8s} {os.path.basename(path)[14:-3]:18s} ctas={ctas}  {t:7.4f} ms  "
                  f"{n * bph / t / 1e6:7.1f} GB/s  same_bits={same}", flush=True)
    del src, tar, H, Href
    torch.cuda.empty_cache()
