"""Ensemble sharding over the GPUs of one box.

Simulations (ensemble members / parameter particles) are independent given the hospital topology -- the
reference's outer loop over simulation columns (/root/reference/src/amro/abm_wards.cpp:92) has no cross-
simulation term -- so the path shards with NO data-path collective: every rank compiles the (replicated)
topology on its own GPU and runs a contiguous block of simulations.  Philox counters use the GLOBAL
simulation index, so results do not depend on how many GPUs share the ensemble.  The only exchange is
one small reduction of the per-day summary (sum and sum of squares of the colonised / detected counts
over simulations).  Two ways to do it: ``PeerMoments`` -- the kernel that computes a run's moments stores them
straight into every GPU's buffer over NVLink (peer memory from torch's symmetric-memory rendezvous; no rank waits
for another inside a run, readers add the ranks' slots up after the barrier that ends a batch of runs) -- and
``reduce_moments``: ``torch.distributed.all_reduce`` over NCCL on GPU tensors, gloo on CPU tensors in the tests.
"""
from __future__ import annotations

import math

SIM_ALIGN = 32  # simulations that share the plane words of the daily draws (rng.cuh); shard boundaries stay on this grid


def shard_sims(total_sims: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous block ``(offset, count)`` of rank ``rank``; offsets are multiples of 32 and the blocks
    tile ``[0, total_sims)`` exactly (trailing ranks may get 0)."""
    groups = math.ceil(total_sims / SIM_ALIGN)
    per = math.ceil(groups / world_size)
    lo = rank * per * SIM_ALIGN
    hi = min((rank + 1) * per * SIM_ALIGN, total_sims)
    return lo, max(hi - lo, 0)


def day_moments(counts_col, counts_det):
    """Per-day (sum, sum of squares) over this rank's simulations, for both metrics: float64 [n_days x 4].
    ``counts_*`` are torch tensors [n_sims x n_days] (the C ABI's column-major [n_days x n_sims])."""
    import torch
    c = counts_col.to(torch.float64)
    d = counts_det.to(torch.float64)
    return torch.stack([c.sum(0), (c * c).sum(0), d.sum(0), (d * d).sum(0)], dim=1)


def reduce_moments(moments, group=None):
    """The ONE collective of a run: sum the per-rank day moments over the ranks, in place (``moments`` must be a
    contiguous tensor: the C ABI's ``day_moments`` buffer as it is, ``[4 x n_days]``, or its copy)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(moments, op=dist.ReduceOp.SUM, group=group)
    return moments


def moments_to_collapse(m, total_sims: int):
    """Reduced day moments ``[n_days x 4]`` -> the columns of ``simulate_and_collapse``: mean colonised, mean detected,
    population sd colonised, population sd detected (reference: abm_wards.cpp:617-620, ``stddev(., 1, 1)``)."""
    import torch
    n = float(total_sims)
    mean = m[:, [0, 2]] / n
    var = torch.clamp(m[:, [1, 3]] / n - mean * mean, min=0.0)
    return torch.cat([mean, var.sqrt()], dim=1)


def collapse(moments, total_sims: int, group=None):
    """``moments`` is float64 [n_days x 4] (``day_moments`` above, or the transpose of the C ABI's
    ``amro_sim_args.day_moments`` buffer, which the fused run fills on the device).
    All-reduce the per-rank day moments (the ONE collective of a run) and turn them into the columns of
    ``simulate_and_collapse``.  Returns float64 [n_days x 4]."""
    import torch
    m = moments.clone(memory_format=torch.contiguous_format)   # a transposed view of the C ABI buffer is not contiguous
    return moments_to_collapse(reduce_moments(m, group), total_sims)


class PeerMoments:
    """The run's collective as peer stores (C ABI: ``amro_day_moments_peers``).

    Every rank owns a buffer ``[n_slots x world x 4 x n_days]`` float64 allocated from torch's symmetric memory and mapped into
    every other rank of the box; ``publish(k, ...)`` enqueues ONE small kernel on the caller's stream that computes this rank's
    per-day moments from its device counts and stores them into slot ``(k % n_slots, rank)`` of every rank's buffer.  An
    NCCL all-reduce per run makes every GPU wait for the slowest one at every run (measured on 8 B200: 0.2-0.9 ms per 4 ms
    run); here nothing waits until ``reduced``: barrier, then the sum over the ranks' slots -- the same numbers as
    ``all_reduce(SUM)`` of the per-rank moments, added in rank order on every rank.  Raises if the ranks cannot map each
    other's memory (callers fall back to ``reduce_moments``)."""

    def __init__(self, n_days: int, device, group=None, n_slots: int = 2):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.n_days, self.n_slots, self.device = int(n_days), int(n_slots), device
        if self.world > 16:
            raise ValueError("PeerMoments: at most 16 ranks")
        self.buf = symm_mem.empty((self.n_slots, self.world, 4, self.n_days), dtype=torch.float64, device=device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        torch.cuda.synchronize(device)
        dist.barrier(self.group)

    def publish(self, k: int, counts_col_ptr: int, counts_det_ptr: int, n_sims: int, ld_counts: int, stream: int = 0):
        """Moments of the run whose device counts are at ``counts_*_ptr`` ([n_days x n_sims] column-major int32) -> slot ``k``."""
        import ctypes as C
        from . import capi
        off = ((k % self.n_slots) * self.world + self.rank) * 4 * self.n_days * 8
        dst = (C.c_void_p * self.world)(*[p + off for p in self.ptrs])
        dev = self.device.index if hasattr(self.device, "index") else int(self.device)
        capi.check(capi.lib().amro_day_moments_peers(counts_col_ptr, counts_det_ptr, self.n_days, n_sims, ld_counts, dst, self.world,
                                                     dev, stream))

    def reduced(self, k: int):
        """float64 ``[4 x n_days]``: the moments of run ``k`` summed over the ranks (every rank gets the same bits)."""
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize(self.device)     # this rank's stores have left ...
        dist.barrier(self.group)                # ... and so have everybody else's
        torch.cuda.synchronize(self.device)
        return self.buf[k % self.n_slots].sum(0)
