"""Keyframe sharding across GPUs: one process per GPU, no data-path collective.

Every (candidate, keyframe) pair is independent up to the final sums
(iba_global.cpp:243-250,322-325,274-275), and everything a keyframe needs from its covisible
keyframes is baked into its own pack rows, so a contiguous block of keyframes per rank needs no
halo.  The only exchange is ONE fp64 sum all-reduce of the per-candidate record
([B,12] for the Nomad path, [B,61] for the LM path); counters travel as doubles (exact < 2^53).
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is plumbing only.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_kf: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous keyframe block of `rank`: sizes differ by at most one."""
    base, rem = divmod(n_kf, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def candidate_tiles(B: int, chunk: int):
    """Candidate chunks processed one after another on every rank (2-D tiling: chunk x shard)."""
    return [(b, min(b + chunk, B)) for b in range(0, B, chunk)]


class SumAllReduce:
    """Callable that all-reduces a host [B, n] fp64 array over the default process group."""

    def __init__(self, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.device = device
        self.enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def __call__(self, arr: np.ndarray) -> np.ndarray:
        if not self.enabled:
            return arr
        t = self.torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64))
        if self.device is not None:
            t = t.to(self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def device_tensor(self, t):
        """In-place all-reduce of a device tensor (NCCL), no host round trip."""
        if self.enabled:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t
