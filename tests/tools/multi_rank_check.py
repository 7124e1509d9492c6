#!/usr/bin/env python
"""Multi-rank GPU check, launched by tests/test_gpu_multi.py under torchrun (one process per
GPU, NCCL): every multi-rank product path against the unsharded single-GPU result."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from sks_homography_b200 import api  # noqa: E402
from sks_homography_b200 import dist as sd  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

P, n_pts, n_hyp, seed, thr2 = 16, 2048, 4096, 7, 2.25
corr = api.synth_corr(P, n_pts, seed=seed, device=dev)
full = api.ransac_keys(corr, n_hyp, seed, thr2)                       # unsharded, on this GPU

# (A) hypothesis shards + NCCL max all-reduce
H, cnt, hyp, _ = sd.ransac_aca(corr, n_hyp, seed, thr2)
kc, kh = api.decode_keys(full)
assert torch.equal(cnt.long(), kc) and torch.equal(hyp, kh), "NCCL-merged winners differ"

# (A') the same, winners merged by the NVLink PeerReducer, several epochs
red = sd.PeerReducer(P, dev)
for _ in range(5):
    H2, cnt2, hyp2, _ = sd.ransac_aca(corr, n_hyp, seed, thr2, reducer=red)
    assert torch.equal(cnt2, cnt) and torch.equal(hyp2, hyp) and torch.equal(H2.view(torch.int32), H.view(torch.int32))
assert not red.timed_out()
red.close()

# (B) pair shards, no collective: rows of the unsharded run
pb, pc = sd.shard_range(P, rank, world)
Hb, cb, hb, _ = sd.ransac_aca_pairs(corr[pb:pb + pc].contiguous(), pb, n_hyp, seed, thr2)
assert torch.equal(cb, cnt[pb:pb + pc]) and torch.equal(hb, hyp[pb:pb + pc])

# streaming solver: contiguous shards of the global index space reassemble the unsharded batch
n = 99_996                      # divisible by 2, 3 and 4: equal shards for all_gather
b, c = sd.shard_range(n, rank, world)
src, tar = api.synth_quads(c, seed=3, dist=1, dtype=torch.float32, device=dev, begin=b)
Hs = api.solve("aca", src, tar)
parts = [torch.empty((sd.shard_range(n, r, world)[1], 9), dtype=torch.float32, device=dev) for r in range(world)]
dist.all_gather(parts, Hs)
s_all, t_all = api.synth_quads(n, seed=3, dist=1, dtype=torch.float32, device=dev)
want = api.solve("aca", s_all, t_all)
got = torch.cat(parts)
assert torch.equal(got.view(torch.int32), want.view(torch.int32)), "sharded batch != unsharded batch"

dist.barrier()
if rank == 0:
    print("MULTI_RANK_OK", world)
dist.destroy_process_group()
