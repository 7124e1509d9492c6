#!/usr/bin/env python
"""The reference's Table 8 workload (imgs/GPU-runtime.png; harness GPU.cu:1166-1243):
time per batch of N = 1 ... 10^6 homographies, fp64 SoA un-normalised, for the
reference's own kernels (oracle/_ref/libsks_refgpu.so, block 32) and ours, with
plain launches and with the launch captured in a CUDA graph (launch-bound regime)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import RefGpuLib  # noqa: E402  (comparator only)
from sks_homography_b200 import api  # noqa: E402

dev = torch.device("cuda:0")
ref = RefGpuLib()


def time_us(fn, iters=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / iters


def graph_us(fn, iters=200, per_graph=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(per_graph):
                fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters // per_graph):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / (iters // per_graph * per_graph)


print(f"{'N':>9s} | {'solver':6s} | {'reference us':>12s} | {'ours us':>9s} | {'ours, CUDA graph us':>19s}")
for n in (1, 10, 100, 1000, 10_000, 100_000, 1_000_000):
    src, tar = api.synth_quads(n, 11, 1, torch.float64, dev, layout="soa")
    H = torch.empty((9, n), dtype=torch.float64, device=dev)
    for solver in ("aca", "sks", "ge", "gpt"):      # the paper's Table 8 also times RHO-GE and GPT-LU
        st = lambda: torch.cuda.current_stream().cuda_stream
        t_ref = time_us(lambda: ref.run(solver, src.data_ptr(), tar.data_ptr(), H.data_ptr(), n, st()))
        ours = lambda: api.solve(solver, src, tar, result=H, normalize=False, layout="soa")
        t_our = time_us(ours)
        t_g = graph_us(ours)
        print(f"{n:9d} | {solver:6s} | {t_ref:12.2f} | {t_our:9.2f} | {t_g:19.2f}", flush=True)


# The reference's whole GPU flow for one batch: get_rand_list + cal_Homo (GPU.cu:1449-1464)
# vs the fused gather+solve kernel, N = 2^24 hypotheses from a 2540-match pool, fp64 SoA.
import numpy as np
n = 1 << 24
pool = torch.from_numpy(np.random.default_rng(0).uniform(7, 790, size=(2540, 4))).to(dev)
# the reference's own sample list: cuRAND host-API MRG32K3A, seed 11 (GPU.cu:1443-1446), from our kernel
t_list = time_us(lambda: api.curand_mrg32k3a(4 * n, 11, dev), 20)
rand4 = api.curand_mrg32k3a(4 * n, 11, dev).view(4, n)
print(f"sample list: 4 x 2^24 MRG32K3A draws (cuRAND host-API order) {t_list:9.1f} us")
H = torch.empty((9, n), dtype=torch.float64, device=dev)
for solver in ("aca", "sks"):
    def two_kernels():
        s, t = api.gather_samples(pool, n, rand4=rand4, layout="soa")
        api.solve(solver, s, t, result=H, normalize=False, layout="soa")
    t2 = time_us(two_kernels, 20)
    t1 = time_us(lambda: api.gather_solve(solver, pool, n, rand4=rand4, normalize=False, layout="soa"), 20)
    print(f"N=2^24 {solver}: gather kernel + solver kernel {t2:9.1f} us | fused gather+solve {t1:9.1f} us  ({t2 / t1:.2f}x)")
