#!/usr/bin/env python
"""Bounded probe of the multi-rank host-path collapse (per-link H2D rate falls
from ~50 GB/s at 1 rank to ~12 GB/s at 8).  One process per GPU under torchrun; every rank copies a
1 GiB host buffer to its GPU in a loop and reports GB/s:

  alone        ranks take turns (one link busy at a time)
  together     all ranks at once, cudaHostAlloc'd buffers (what bench.py's e2e leg uses)
  hugepages    all at once, anonymous mmap + MADV_HUGEPAGE + cudaHostRegister
  staggered    all at once, rank r started r x 25 ms late (DMA bursts de-phased)
  bidir        all at once, H2D and D2H on two streams (the e2e pipeline's pattern)

Record and stop: this is a measurement of the box, not a tuning loop.
    torchrun --nproc-per-node N tools/pcie_multi_probe.py"""
import ctypes
import json
import mmap
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N = 1 << 30
d = torch.empty(N, dtype=torch.uint8, device=dev)
d2 = torch.empty(N, dtype=torch.uint8, device=dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def rate(copy, seconds=0.6, delay=0.0):
    barrier()
    if delay:
        time.sleep(delay)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    k = 0
    while time.perf_counter() - t0 < seconds:
        copy()
        k += 1
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return k * N / e0.elapsed_time(e1) / 1e6


def gather(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    out = [torch.zeros_like(t) for _ in range(world)]
    if world > 1:
        dist.all_gather(out, t)
    else:
        out = [t]
    return [round(float(o.item()), 3) for o in out]


res = {"world": world, "buffer_GiB": 1}
h = torch.empty(N, dtype=torch.uint8, pin_memory=True)
h2 = torch.empty(N, dtype=torch.uint8, pin_memory=True)
h.fill_(1)
alone = []
for r in range(world):
    barrier()
    v = rate(lambda: d.copy_(h, non_blocking=True), 0.3) if r == rank else 0.0
    if r != rank:
        barrier()                                   # matches the barrier inside rate()
    alone.append(v)
res["alone_GBps"] = gather(max(alone))
res["together_GBps"] = gather(rate(lambda: d.copy_(h, non_blocking=True)))
res["staggered_GBps"] = gather(rate(lambda: d.copy_(h, non_blocking=True), delay=0.025 * rank))
s2 = torch.cuda.Stream()


def bidir():
    d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


res["bidir_h2d_GBps"] = gather(rate(bidir))
# the e2e pipeline's traffic mix (64 B in : 36 B out per homography), two ways: both directions at once on
# two streams (what the pipeline does), or globally ALIGNED phases -- every rank copies in, barrier, every
# rank copies out, barrier -- so that host->device and device->host traffic never share the host
IN, OUT = 64 << 20, 36 << 20
hi, ho = h[:IN], h2[:OUT]
di, do = d[:IN], d2[:OUT]


def mixed(aligned, cycles=24):
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    for _ in range(cycles):
        if aligned:
            di.copy_(hi, non_blocking=True)
            barrier()
            ho.copy_(do, non_blocking=True)
            barrier()
        else:
            di.copy_(hi, non_blocking=True)
            with torch.cuda.stream(s2):
                ho.copy_(do, non_blocking=True)
            torch.cuda.synchronize()
    barrier()
    el = time.perf_counter() - t0
    return cycles * (IN / 64) / el / 1e9          # G homographies/s this rank could feed


res["mixed_overlapped_GHps"] = gather(round(mixed(False), 4))
res["mixed_aligned_GHps"] = gather(round(mixed(True), 4))
d2h_rate = []
s3 = torch.cuda.Stream()


def bidir_d2h():
    with torch.cuda.stream(s3):
        d.copy_(h, non_blocking=True)
    h2.copy_(d2, non_blocking=True)


res["bidir_d2h_GBps"] = gather(rate(bidir_d2h))
res["together_d2h_GBps"] = gather(rate(lambda: h2.copy_(d2, non_blocking=True)))
# transparent-hugepage-backed host memory, registered with CUDA
try:
    mm = mmap.mmap(-1, N, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    mm.madvise(mmap.MADV_HUGEPAGE)
    buf = (ctypes.c_char * N).from_buffer(mm)
    ctypes.memset(ctypes.addressof(buf), 1, N)      # touch: pages are faulted in (2 MiB where THP allows)
    ptr = ctypes.addressof(buf)
    rt = torch.cuda.cudart()
    err = rt.cudaHostRegister(ptr, N, 0)
    ht = torch.frombuffer(mm, dtype=torch.uint8)
    thp = ""
    try:
        thp = open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip()
        anon = [l for l in open("/proc/self/smaps_rollup") if l.startswith("AnonHugePages")]
        thp += " | " + (anon[0].strip() if anon else "")
    except Exception:
        pass
    res["hugepages_GBps"] = gather(rate(lambda: d.copy_(ht, non_blocking=True)))
    res["hugepages_note"] = f"cudaHostRegister -> {err}; THP: {thp}"
    rt.cudaHostUnregister(ptr)
except Exception as e:
    res["hugepages_error"] = f"{type(e).__name__}: {e}"[:200]
    barrier()
if rank == 0:
    try:
        res["host"] = {"cpus": os.cpu_count(), "numa_nodes": len([x for x in os.listdir("/sys/devices/system/node") if x.startswith("node")])}
    except Exception:
        pass
    for k in ("mixed_overlapped_GHps", "mixed_aligned_GHps"):
        res[k.replace("_GHps", "_sum_GHps")] = round(sum(res[k]), 3)
    for k in ("alone_GBps", "together_GBps", "staggered_GBps", "bidir_h2d_GBps", "bidir_d2h_GBps", "together_d2h_GBps",
              "hugepages_GBps"):
        if k in res:
            res[k.replace("_GBps", "_sum_GBps")] = round(sum(res[k]), 1)
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
