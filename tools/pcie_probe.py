#!/usr/bin/env python
"""PCIe probe: H2D bandwidth with 1, 2, 4 concurrent copy streams, D2H alone, and
both directions at once (pinned memory, 256 MiB per copy)."""
import time
import torch

dev = torch.device("cuda:0")
MB = 256
n = MB << 20
host = [torch.empty(n, dtype=torch.uint8, pin_memory=True) for _ in range(4)]
devb = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(4)]
hout = [torch.empty(n, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
streams = [torch.cuda.Stream() for _ in range(6)]


def run(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d(k):
    def f():
        for i in range(k):
            with torch.cuda.stream(streams[i]):
                for _ in range(4 // k):
                    devb[i].copy_(host[i], non_blocking=True)
    return f


for k in (1, 2, 4):
    t = run(h2d(k))
    print(f"H2D {k} stream(s): {4 * n / t / 1e9:.1f} GB/s")


def d2h():
    with torch.cuda.stream(streams[4]):
        for _ in range(2):
            hout[0].copy_(devb[0], non_blocking=True)


t = run(d2h)
print(f"D2H 1 stream: {2 * n / t / 1e9:.1f} GB/s")


def both(k):
    def f():
        h2d(k)()
        with torch.cuda.stream(streams[4]):
            for _ in range(2):
                hout[0].copy_(devb[3], non_blocking=True)
        with torch.cuda.stream(streams[5]):
            hout[1].copy_(devb[2], non_blocking=True) if k > 2 else None
    return f


for k in (1, 2):
    t = run(both(k))
    print(f"bidirectional, H2D on {k} stream(s): H2D {4 * n / t / 1e9:.1f} GB/s + D2H {2 * n / t / 1e9:.1f} GB/s in the same time")
