#!/usr/bin/env python
"""BASELINE configs[0] -- the reference's own "C++ Runtime Test" (CPU/main.cpp:87-114) restated:
10^6 random correspondence quadruples, fp32 (and fp64), SKS / ACA (and the competitor GE) on the
host, timed with a monotonic clock instead of cv::getTickCount (OpenCV C++ is absent).  Two
regimes: the reference's (ONE quadruple re-used for every call, CPU/main.cpp:89-91: pure latency
chain, what the paper's Table 5 reports) and distinct quadruples streamed from memory, on one
thread and on all host threads.  Uses the reference's C++ compiled in place (oracle/_ref):
test/bench infrastructure, not the product."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.oracle import Oracle, RefLib   # noqa: E402

PAPER_US = {("aca", "f32"): 0.0145, ("aca", "f64"): 0.0171, ("sks", "f32"): 0.0252, ("sks", "f64"): 0.0256,
            ("ge", "f32"): 0.0287}      # imgs/CPU-runtime.png, row O2 (i7-10700)
N = 1_000_000
o = Oracle()


def best_of(fn, k=5):
    fn()
    b = 1e30
    for _ in range(k):
        t0 = time.perf_counter(); fn(); b = min(b, time.perf_counter() - t0)
    return b


for o3 in (False, True):
    ref = RefLib(o3=o3)
    T = ref.hardware_threads()
    print(f"== {'g++ -O3 -march=x86-64-v3 -ffp-contract=fast' if o3 else 'g++ -O2 -ffp-contract=off (parity build)'}"
          f", {T} host threads")
    print(f"{'solver':10s} {'same quad, 1 thr':>18s} {'distinct, 1 thr':>18s} {'distinct, all thr':>18s} {'paper O2':>10s}   (us per homography)")
    for dt, tag in ((np.float32, "f32"), (np.float64, "f64")):
        s, t = o.synth_quads(0, N, 11, 1, dt)
        R = 1 << 16                                   # one quadruple repeated, buffers stay cache-resident
        same_s, same_t = np.repeat(s[:1], R, 0), np.repeat(t[:1], R, 0)
        out = np.empty((N, 9), dtype=dt)
        for solver in ("aca", "sks", "ge"):
            if solver == "ge" and tag == "f64":
                continue
            a = best_of(lambda: ref.solve(solver, same_s, same_t, threads=1, out=out[:R]), 20) / R
            b = best_of(lambda: ref.solve(solver, s, t, threads=1, out=out)) / N
            c = best_of(lambda: ref.solve(solver, s, t, threads=T, out=out)) / N
            p = PAPER_US.get((solver, tag))
            print(f"{solver + '_' + tag:10s} {a * 1e6:18.4f} {b * 1e6:18.4f} {c * 1e6:18.4f} {p if p else float('nan'):10.4f}")
