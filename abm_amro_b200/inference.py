"""Parameter-inference driver around the fused kernels.

The reference's tutorial fits the six transmission parameters by calling ``simulate_and_collapse`` in an optimisation loop
(/root/reference/tutorials/Test_2024.ipynb:927).  On the device the natural shape of that loop is a SWEEP: tens of thousands
of parameter particles drawn from a prior are simulated in one launch -- one particle per simulation column, each with its
own parameter row (/root/reference/src/amro/abm_wards.cpp:79-84, 98-112) -- and what comes back is one number per particle:
its squared distance to the observed per-day series (``amro_sim_args.distance``, computed on the device), not the
[days x particles] counts.  An approximate-Bayesian-computation rejection step then keeps the closest fraction.

Particles shard over the GPUs of a box like ensemble members do (``ensemble.shard_sims``): every rank simulates a block of
particles with the GLOBAL particle index in the Philox counters and the ranks' distance vectors are gathered once.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import capi, ensemble

PARAMETERS = ("alpha", "beta", "gamma", "rho_hospital", "alpha2", "rho_imported")   # columns of `parameters` (abm_wards.cpp:79-84)


@dataclass
class UniformPrior:
    """Independent uniform priors, one (low, high) pair per parameter; the defaults are the ranges of SURVEY.md 8(d)."""
    alpha: tuple = (0.05, 0.3)
    beta: tuple = (0.2, 0.9)
    gamma: tuple = (0.02, 0.2)
    rho_hospital: tuple = (0.3, 0.9)
    alpha2: tuple = (0.1, 0.5)
    rho_imported: tuple = (0.5, 1.0)

    def bounds(self) -> np.ndarray:
        return np.array([getattr(self, p) for p in PARAMETERS], dtype=np.float64)      # [6 x 2]

    def sample(self, n: int, seed: int = 100) -> np.ndarray:
        """``n`` particles, float64 [n x 6] in the column order of the reference's ``parameters``."""
        b = self.bounds()
        return np.asfortranarray(b[:, 0] + np.random.default_rng(seed).random((n, 6)) * (b[:, 1] - b[:, 0]))


@dataclass
class SweepResult:
    particles: np.ndarray        # [n x 6] all particles of the sweep (every rank holds all of them)
    distance: np.ndarray         # [n] squared distance of every particle's per-day counts to the data
    accepted: np.ndarray         # indices of the closest ``keep`` fraction, closest first
    posterior_mean: np.ndarray   # [6] mean of the accepted particles
    posterior_sd: np.ndarray     # [6]
    kernel_ms: float             # this rank's fused-kernel time


def abc_sweep(topo, observed_colonized=None, observed_detected=None, *, n_particles: int = 65_536, prior: UniformPrior | None = None,
              initial_colonized_probability: float = 0.2, initial_detected_probability: float = 0.2, time_to_detect: int = 2,
              testing_schedule_hospitalized: int = 7, testing_schedule_arrivals: int = 1, seed: int = 42, prior_seed: int = 100,
              keep: float = 0.01, group=None) -> SweepResult:
    """One sweep of ``n_particles`` prior draws on the compiled hospital ``topo`` (an ``engine.Topology`` on this rank's GPU).

    ``observed_*`` are per-day series of length ``topo.info.n_days_out`` (NaN = no data that day); the schedules are the
    reference's ``testing_schedule_*`` arguments (7 = weekly).  With ``torch.distributed`` initialised the particles are
    split over the ranks and the distances gathered; every rank returns the full result.
    """
    prior = prior or UniformPrior()
    particles = prior.sample(n_particles, prior_seed)
    rank, world = 0, 1
    dist = None
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(group), dist.get_world_size(group)
    except ImportError:
        dist = None
    off, n = ensemble.shard_sims(n_particles, world, rank)
    mine = np.zeros(0)
    kernel_ms = 0.0
    if n > 0:
        r = topo.simulate(particles[off:off + n], initial_colonized_probability, initial_detected_probability, time_to_detect,
                          testing_schedule_hospitalized, testing_schedule_arrivals, seed, sim_offset=off, want_counts=False,
                          observed_colonized=observed_colonized, observed_detected=observed_detected)
        if r.path_used != capi.PATH_FAST:
            raise RuntimeError("the sweep fell off the fused kernels; check topo.info.why_not_fast")
        mine, kernel_ms = r.distance, r.kernel_ms
    if world > 1:
        import torch
        blocks = [ensemble.shard_sims(n_particles, world, k) for k in range(world)]
        width = max(b[1] for b in blocks)
        backend = dist.get_backend(group)
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        send = torch.full((width,), float("inf"), dtype=torch.float64, device=dev)
        send[:n] = torch.from_numpy(np.ascontiguousarray(mine)).to(dev)
        recv = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(recv, send, group=group)                      # the sweep's one collective: 8 bytes per particle
        distance = np.concatenate([t.cpu().numpy()[:b[1]] for t, b in zip(recv, blocks)])
    else:
        distance = np.asarray(mine)
    n_keep = max(1, int(round(keep * n_particles)))
    accepted = np.argsort(distance, kind="stable")[:n_keep]
    post = particles[accepted]
    return SweepResult(particles, distance, accepted, post.mean(0), post.std(0), kernel_ms)
