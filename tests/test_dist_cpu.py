"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: contiguous sharding
of the streaming batch and the one max-all-reduce that merges the RANSAC winners.
The per-rank compute here is the oracle standing in for the GPU kernel; the code
under test is sks_homography_b200.dist (shard_range / merge_keys)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.oracle import Oracle
        from sks_homography_b200 import dist as sd
        o = Oracle()
        # --- streaming path: contiguous shards, no collective
        n = 10_007
        b, c = sd.shard_range(n)
        s, t = o.synth_quads(b, c, 11, 1, np.float32)
        H = o.solve("aca", s, t)
        np.save(os.path.join(out_dir, f"H_{rank}.npy"), H)
        np.save(os.path.join(out_dir, f"range_{rank}.npy"), np.array([b, c]))
        # --- RANSAC path: hypothesis shards + one integer max all-reduce
        rng = np.random.default_rng(3)
        corr = rng.uniform(10, 150, size=(4, 128, 4)).astype(np.float32)
        corr[:, :80, 2:] = corr[:, :80, :2] * 1.05 + 3.0          # a planted similarity
        n_hyp = 203
        hb, hc = sd.shard_range(n_hyp)
        keys = o.ransac(corr, n_hyp, seed=17, thr2=4.0, hyp_begin=hb, hyp_count=hc)
        kt = torch.from_numpy(keys.view(np.int64).copy())
        sd.merge_keys(kt)
        np.save(os.path.join(out_dir, f"keys_{rank}.npy"), kt.numpy())
        # --- RANSAC variant B: image pairs sharded, no collective; the global pair id keys the sampler
        pb, pc = sd.shard_range(corr.shape[0])
        kp = o.ransac(corr[pb:pb + pc], n_hyp, seed=17, thr2=4.0, pair_begin=pb)
        np.save(os.path.join(out_dir, f"pairkeys_{rank}.npy"), kp)
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_key_merge(tmp_path, oracle):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    # shards tile [0, n) and their concatenation is the single-process result, bytewise
    ranges = [np.load(tmp_path / f"range_{r}.npy") for r in range(world)]
    assert ranges[0][0] == 0 and ranges[1][0] == ranges[0][1] and ranges[1].sum() == 10_007
    H = np.concatenate([np.load(tmp_path / f"H_{r}.npy") for r in range(world)])
    s, t = oracle.synth_quads(0, 10_007, 11, 1, np.float32)
    assert np.array_equal(H.view(np.uint32), oracle.solve("aca", s, t).view(np.uint32))
    # merged keys: identical on both ranks and equal to the unsharded oracle
    rng = np.random.default_rng(3)
    corr = rng.uniform(10, 150, size=(4, 128, 4)).astype(np.float32)
    corr[:, :80, 2:] = corr[:, :80, :2] * 1.05 + 3.0
    full = oracle.ransac(corr, 203, seed=17, thr2=4.0)
    k0, k1 = (np.load(tmp_path / f"keys_{r}.npy") for r in range(world))
    assert np.array_equal(k0, k1)
    assert np.array_equal(k0.view(np.uint64), full)
    assert ((full >> np.uint64(32)) >= 60).all()
    # pair shards: concatenated == unsharded, without any exchange
    assert np.array_equal(np.concatenate([np.load(tmp_path / f"pairkeys_{r}.npy") for r in range(world)]), full)
