"""Multi-GPU driver: one process per GPU, torch.distributed for the plumbing.

Streaming solvers shard the batch into contiguous index ranges and need no
collective at all; fused ACA-RANSAC shards the hypothesis ids
of every image pair and merges the per-rank best keys with ONE integer
max-all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests).  The winning
model is recomputed from its id on every rank, so no homography travels.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from ._lib import lib


def world(group=None) -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(n: int, rank: int | None = None, world_size: int | None = None, group=None
                ) -> tuple[int, int]:
    """(begin, count) of this rank's contiguous shard of n units."""
    r, w = world(group)
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    return lib().shard_range(n, rank, world_size)


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pin this process's host threads to the CPUs of the NUMA node its GPU hangs off, so that
    the pinned host buffers it allocates afterwards (first touch) and the staging memcpys of
    the host-pointer path stay on the socket that owns the PCIe link.  One process per GPU on
    a two-socket box otherwise puts about half of the ranks' buffers across the socket
    interconnect.  Pure sysfs + sched_setaffinity; a no-op (with the reason reported) where
    the topology is not visible."""
    import os
    info = {"bound": False}
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        node = int(open(f"{base}/numa_node").read())
        cpulist = open(f"{base}/local_cpulist").read().strip()
        info.update(pci=bdf, numa_node=node, local_cpulist=cpulist)
        cpus = set()
        for part in cpulist.split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if node < 0 or not cpus or cpus == allowed:
            info["reason"] = "single node or topology not exposed"
            return info
        os.sched_setaffinity(0, cpus)
        info.update(bound=True, cpus=len(cpus))
    except Exception as e:                                    # sysfs absent, container limits...
        info["reason"] = f"{type(e).__name__}: {e}"[:120]
    return info


def merge_keys(keys: torch.Tensor, group=None) -> torch.Tensor:
    """In-place max-combine of packed (count<<32 | ~hyp) keys across ranks.

    The keys are < 2^63 (counts are < 2^31), so the signed int64 max NCCL / gloo
    implement orders them exactly like the unsigned keys.
    """
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(keys, op=dist.ReduceOp.MAX, group=group)
    return keys


class PeerReducer:
    """Hand-written NVLink replacement for the one NCCL all-reduce of the RANSAC step
    (csrc/peer.cuh): every rank's keys are max-combined into every rank's exchange
    block with system-scope atomics over peer memory; two small kernels per rank, no
    collective call inside the step.  Setup (once) exchanges CUDA-IPC handles through
    torch.distributed.  Works for world == 1 too (the block is then only the own one)."""

    def __init__(self, n_keys: int, device, group=None, timeout_s: float = 5.0):
        import ctypes as C
        self.L = lib()
        self.group = group
        self.rank, self.world = world(group)
        self.n_keys = int(n_keys)
        self.device = torch.device(device)
        self.timeout_s = float(timeout_s)
        self.epoch = 0
        self._opened = []
        with torch.cuda.device(self.device):
            own = C.c_void_p()
            self.L.check(self.L.c.sks_cuda_peer_alloc(C.byref(own), self.n_keys), "sks_cuda_peer_alloc")
            self.own = own
            handle = (C.c_ubyte * 64)()
            self.L.check(self.L.c.sks_cuda_peer_export(own, handle), "sks_cuda_peer_export")
            handles = [bytes(handle)]
            if self.world > 1:
                handles = [None] * self.world
                dist.all_gather_object(handles, bytes(handle), group=group)
            self.blocks = (C.c_void_p * self.world)()
            for g in range(self.world):
                if g == self.rank:
                    self.blocks[g] = own
                else:
                    p = C.c_void_p()
                    buf = (C.c_ubyte * 64).from_buffer_copy(handles[g])
                    self.L.check(self.L.c.sks_cuda_peer_open(buf, C.byref(p)), "sks_cuda_peer_open")
                    self.blocks[g] = p
                    self._opened.append(p)
            # sticky time-out flag in pinned host memory (device-visible under UVA): the wait kernel
            # sets it to 1, the host reads it without synchronising the stream
            self.status = torch.zeros(1, dtype=torch.int32).pin_memory()
            if self.world > 1:
                dist.barrier(group=group)      # every block exists and is zeroed before the first push

    def max_reduce_(self, keys: torch.Tensor) -> torch.Tensor:
        """In-place global max of the packed keys (int64 view) across the ranks."""
        if keys.numel() != self.n_keys or keys.dtype != torch.int64 or not keys.is_contiguous():
            raise ValueError("keys must be a contiguous int64 tensor of n_keys elements")
        self.check()                   # a time-out of an EARLIER step is fatal for the exchange
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream(self.device).cuda_stream
            self.L.check(self.L.c.sks_cuda_peer_push_max(keys.data_ptr(), self.n_keys, self.blocks, self.world,
                                                         self.rank, self.epoch, st), "sks_cuda_peer_push_max")
            self.L.check(self.L.c.sks_cuda_peer_wait(self.own, self.world, self.epoch, keys.data_ptr(),
                                                     self.n_keys, self.status.data_ptr(), self.timeout_s, st),
                         "sks_cuda_peer_wait")
        self.epoch += 1
        return keys

    def timed_out(self) -> bool:
        """Synchronises the device, then reports whether any wait so far timed out."""
        torch.cuda.synchronize(self.device)
        return bool(self.status[0].item())

    def check(self) -> None:
        """Raise if a wait timed out (no synchronisation: sees every step that has completed).
        After a time-out the keys of that step were left un-reduced and the arrival counters no
        longer line up, so the reducer cannot be used again: rebuild it or use merge_keys()."""
        if int(self.status[0]) != 0:
            raise RuntimeError(
                f"PeerReducer (rank {self.rank}/{self.world}): a peer did not arrive within "
                f"{self.timeout_s} s; the keys of that step are NOT reduced and the exchange is "
                "unusable -- rebuild the reducer or fall back to dist.merge_keys()")

    def close(self, check: bool = True) -> None:
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            failed = int(self.status[0]) != 0
            if self.world > 1:
                dist.barrier(group=self.group)
            for p in self._opened:
                self.L.c.sks_cuda_peer_close(p)
            self._opened = []
            if self.own is not None:
                self.L.c.sks_cuda_peer_free(self.own)
                self.own = None
        if check and failed:
            raise RuntimeError("PeerReducer: a wait timed out during this session (keys of that step un-reduced)")


def ransac_aca(corr: torch.Tensor, n_hyp: int, seed: int, thr2: float,
               samples: torch.Tensor | None = None, group=None, want_mask: bool = False,
               reducer: "PeerReducer | None" = None):
    """Hypothesis-sharded fused ACA-RANSAC.  Every rank holds all pairs' matches
    `corr` [P, n_pts, 4]; returns (H_best [P,9], inlier_count [P], hyp_id [P], mask)."""
    from . import api
    begin, count = shard_range(n_hyp, group=group)
    keys = api.ransac_keys(corr, n_hyp, seed, thr2, samples, begin, count)
    if reducer is not None:
        reducer.max_reduce_(keys)      # NVLink peer atomics instead of the NCCL all-reduce
        if reducer.timed_out():        # synchronises; the models below would otherwise differ per rank
            reducer.check()
    else:
        merge_keys(keys, group)
    H, cnt, mask = api.ransac_finalize(corr, n_hyp, seed, thr2, keys, samples, want_mask)
    _, hyp = api.decode_keys(keys)
    return H, cnt, hyp, mask


def ransac_aca_pairs(corr_shard: torch.Tensor, pair_begin: int, n_hyp: int, seed: int, thr2: float,
                     samples: torch.Tensor | None = None, want_mask: bool = False):
    """Pair-sharded fused ACA-RANSAC (SURVEY.md 8(e) variant B): this rank owns image pairs
    [pair_begin, pair_begin + P_local) and scores ALL hypothesis ids of them; no collective.
    `pair_begin` keys the sampler, so the result equals rows [pair_begin, ...) of an unsharded run.
    Returns (H_best [P_local,9], inlier_count, hyp_id, mask)."""
    from . import api
    keys = api.ransac_keys(corr_shard, n_hyp, seed, thr2, samples, pair_begin=pair_begin)
    H, cnt, mask = api.ransac_finalize(corr_shard, n_hyp, seed, thr2, keys, samples, want_mask,
                                       pair_begin=pair_begin)
    _, hyp = api.decode_keys(keys)
    return H, cnt, hyp, mask
