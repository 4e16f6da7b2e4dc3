"""Ensemble sharding over the GPUs of one box.

Simulations (ensemble members / parameter particles) are independent given the hospital topology -- the
reference's outer loop over simulation columns (/root/reference/src/amro/abm_wards.cpp:92) has no cross-
simulation term -- so the path shards with NO data-path collective: every rank compiles the (replicated)
topology on its own GPU and runs a contiguous block of simulations.  Philox counters use the GLOBAL
simulation index, so results do not depend on how many GPUs share the ensemble.  The only exchange is
one small reduction of the per-day summary (sum and sum of squares of the colonised / detected counts
over simulations): ``torch.distributed.all_reduce`` over NCCL/NVLink on GPU tensors, gloo on CPU tensors
in the tests.
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
