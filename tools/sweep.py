#!/usr/bin/env python
"""In-process tuning sweep of the streaming kernels (one input allocation per
workload, CUDA-event timing, inputs >> L2).  Prints one line per configuration:
achieved algorithmic GB/s and fraction of the measured HBM copy peak."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS, peaks  # noqa: E402
from sks_homography_b200 import api, lib  # noqa: E402


def time_ms(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="aca_f32,sks_f32,rect_f32,aca_f64,sks_f64")
    ap.add_argument("--log2n", type=int, default=None)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    L = lib()
    peak, _ = peaks()
    dev = torch.device("cuda:0")
    configs = [("direct", 1, 0, 4, 0)]
    for small in (0, 1):
        for stages in (2, 3, 4, 6, 8):
            for ctas in (0, 1, 2):
                configs.append((f"ring small={small} stages={stages} ctas={ctas}", 2, small, stages, ctas))
    if args.quick:
        configs = [c for c in configs if c[1] == 1 or (c[3] in (3, 4) and c[4] == 0)]
    rows = []
    for w in args.workloads.split(","):
        solver, dt, bph, log2n, dist = WORKLOADS[w]
        n = 1 << (args.log2n or log2n)
        tdt = torch.float32 if dt == "f32" else torch.float64
        src, tar = api.synth_quads(n, 11, dist, tdt, dev)
        H = torch.empty((n, 9), dtype=tdt, device=dev)
        if solver == "rect":
            run = lambda: api.aca_rect(tar, 128.0, 1.0, 15.0, 12.0, result=H)
        else:
            run = lambda: api.solve(solver, src, tar, result=H)
        for name, variant, small, stages, ctas in configs:
            L.c.sks_cuda_set_variant(variant)
            L.c.sks_cuda_set_tuning(small, stages, ctas)
            med, best = time_ms(run)
            gbs = n * bph / med / 1e6
            rows.append({"workload": w, "config": name, "ms": med, "ms_min": best, "GBps": gbs,
                         "frac": gbs / peak, "GHps": n / med / 1e6})
            print(f"{w:9s} {name:34s} {med:8.3f} ms  {gbs:8.1f} GB/s  {gbs / peak:6.3f} of peak  "
                  f"{n / med / 1e6:7.2f} GH/s", flush=True)
        # SoA (reference GPU layout)
        L.c.sks_cuda_set_variant(0)
        L.c.sks_cuda_set_tuning(0, 4, 0)
        s2, t2 = src.T.contiguous(), tar.T.contiguous()
        del src, tar
        H2 = H.view(9, n)
        if solver == "rect":
            run = lambda: api.aca_rect(t2, 128.0, 1.0, 15.0, 12.0, result=H2, layout="soa")
        else:
            run = lambda: api.solve(solver, s2, t2, result=H2, layout="soa")
        med, best = time_ms(run)
        gbs = n * bph / med / 1e6
        rows.append({"workload": w, "config": "soa", "ms": med, "ms_min": best, "GBps": gbs,
                     "frac": gbs / peak, "GHps": n / med / 1e6})
        print(f"{w:9s} {'soa':34s} {med:8.3f} ms  {gbs:8.1f} GB/s  {gbs / peak:6.3f} of peak  "
              f"{n / med / 1e6:7.2f} GH/s", flush=True)
        del s2, t2, H, H2
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
