#!/usr/bin/env python
"""torchrun test of the NVLink peer max-reduce against the NCCL all-reduce:
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/peer_reduce_probe.py"""
import os, sys, time
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sks_homography_b200 import api, dist as sd  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
P = 1024
red = sd.PeerReducer(P, dev)
ok = True
for epoch in range(20):
    g = torch.Generator().manual_seed(1000 * epoch + rank)
    keys = torch.randint(0, 2**62, (P,), generator=g, dtype=torch.int64).to(dev)
    ref = keys.clone()
    dist.all_reduce(ref, op=dist.ReduceOp.MAX)
    red.max_reduce_(keys)
    ok = ok and bool(torch.equal(keys, ref)) and not red.timed_out()
# full RANSAC step both ways
corr = api.synth_corr(64, 4096, seed=11, device=dev)
a = sd.ransac_aca(corr, 8192, 11, 2.25)
b = sd.ransac_aca(corr, 8192, 11, 2.25, reducer=sd.PeerReducer(64, dev))
ok = ok and all(torch.equal(x, y) for x, y in zip(a[:3], b[:3]))
# latency of the two reduces
def lat(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return 1e6 * (time.perf_counter() - t0) / n
keys = torch.zeros(P, dtype=torch.int64, device=dev)
t_nccl = lat(lambda: dist.all_reduce(keys, op=dist.ReduceOp.MAX))
t_peer = lat(lambda: red.max_reduce_(keys))
okt = torch.tensor([int(ok)], device=dev); dist.all_reduce(okt, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world={world} peer max-reduce == NCCL all-reduce over 20 epochs and a full RANSAC step: {bool(okt.item())}")
    print(f"latency per 8 KiB reduce: NCCL {t_nccl:.1f} us | NVLink peer atomics {t_peer:.1f} us")
red.close()
dist.destroy_process_group()
