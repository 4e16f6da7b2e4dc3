"""Synthetic hospitals of the shapes named in BASELINE.json (SURVEY.md section 8(d) generator).

The reference ships no data (its tutorial's CSVs are git-ignored, /root/reference/.gitignore:5,17-18);
the tables built here have exactly the layout its entry points take
(/root/reference/src/amro/abm_wards.cpp:87-90,291-295,335-338):

  ward_matrix [R x 6] float64, order='F':  day | ward | MRN | is_new | weight | next_day
  total_patients_per_ward [K x 3]:          day | ward | N (= sum of weights of that ward-day)

Rows are emitted day by day, sorted by (ward, MRN) inside a day; ``next_day`` is the row index of the
same patient's record on the following day (-1 when discharged).
"""
from __future__ import annotations

import numpy as np

TUTORIAL_PARAMS = dict(alpha=0.15, beta=0.55, gamma=0.05, rho_hospital=0.6, alpha2=0.35, rho_imported=0.9)


def make_hospital(n_wards: int, census: int, n_days: int, seed: int = 1,
                  p_discharge: float = 0.15, p_transfer: float = 0.10, p_split: float = 0.0):
    """Return ``(ward_matrix, total_patients_per_ward, n_init)``.

    Every day each of the ``census`` beds emits one patient-day record (or, with probability
    ``p_split``, two records in two different wards with weight 0.5 each).  Afterwards the patient is
    discharged with probability ``p_discharge`` and replaced by a new MRN in a uniformly random ward
    (``is_new`` = 1 on its first day), otherwise transferred to a uniformly random ward with probability
    ``p_transfer``.
    """
    rng = np.random.default_rng(seed)
    mrn = np.arange(1, census + 1, dtype=np.int64)
    next_mrn = census + 1
    ward = rng.integers(1, n_wards + 1, size=census)
    is_new = np.ones(census, dtype=np.int64)
    same_as_yesterday = np.zeros(census, dtype=bool)
    prev_rows = np.full((census, 2), -1, dtype=np.int64)

    days, wards, mrns, news, weights = [], [], [], [], []
    next_day = []
    tppw = []
    offset = 0
    for d in range(n_days):
        if p_split > 0.0 and n_wards >= 2:
            split = rng.random(census) < p_split
            other = (ward - 1 + rng.integers(1, n_wards, size=census)) % n_wards + 1
        else:
            split = np.zeros(census, dtype=bool)
            other = ward
        slot = np.concatenate([np.arange(census), np.nonzero(split)[0]])
        k = np.concatenate([np.zeros(census, dtype=np.int64), np.ones(int(split.sum()), dtype=np.int64)])
        r_ward = np.concatenate([ward, other[split]])
        r_w = np.where(split[slot], 0.5, 1.0)
        order = np.lexsort((k, mrn[slot], r_ward))
        slot, k, r_ward, r_w = slot[order], k[order], r_ward[order], r_w[order]
        n = slot.size
        gidx = offset + np.arange(n, dtype=np.int64)
        today_rows = np.full((census, 2), -1, dtype=np.int64)
        today_rows[slot, k] = gidx
        last = np.where(today_rows[:, 1] >= 0, 1, 0)

        # link yesterday's records to today's
        if d > 0:
            nd_prev = next_day[-1]
            for kk in (0, 1):
                src = prev_rows[:, kk]
                ok = (src >= 0) & same_as_yesterday
                tgt = today_rows[np.arange(census), np.minimum(kk, last)]
                nd_prev[src[ok] - prev_offset] = tgt[ok]

        days.append(np.full(n, d, dtype=np.float64))
        wards.append(r_ward.astype(np.float64))
        mrns.append(mrn[slot].astype(np.float64))
        news.append(is_new[slot].astype(np.float64))
        weights.append(r_w.astype(np.float64))
        next_day.append(np.full(n, -1, dtype=np.int64))

        tot = np.bincount(r_ward, weights=r_w, minlength=n_wards + 1)
        present = np.nonzero(np.bincount(r_ward, minlength=n_wards + 1))[0]
        tppw.append(np.stack([np.full(present.size, d, dtype=np.float64), present.astype(np.float64), tot[present]], axis=1))

        prev_rows, prev_offset = today_rows, offset
        offset += n

        # transitions for tomorrow
        disc = rng.random(census) < p_discharge
        trans = (~disc) & (rng.random(census) < p_transfer)
        new_w = rng.integers(1, n_wards + 1, size=census)
        nd = int(disc.sum())
        mrn = mrn.copy()
        mrn[disc] = np.arange(next_mrn, next_mrn + nd, dtype=np.int64)
        next_mrn += nd
        ward = np.where(disc | trans, new_w, ward)
        is_new = disc.astype(np.int64)
        same_as_yesterday = ~disc

    R = offset
    wm = np.empty((R, 6), dtype=np.float64, order="F")
    wm[:, 0] = np.concatenate(days)
    wm[:, 1] = np.concatenate(wards)
    wm[:, 2] = np.concatenate(mrns)
    wm[:, 3] = np.concatenate(news)
    wm[:, 4] = np.concatenate(weights)
    wm[:, 5] = np.concatenate(next_day).astype(np.float64)
    tp = np.asfortranarray(np.concatenate(tppw, axis=0))
    n_init = int(days[0].size)
    return wm, tp, n_init


def tutorial_parameters(n_sims: int) -> np.ndarray:
    """``parameters[S x 6]`` (alpha, beta, gamma, rho_hospital, alpha2, rho_imported), every row the
    tutorial's values (/root/reference/tutorials/Test_2024.ipynb cells 13/21)."""
    p = TUTORIAL_PARAMS
    row = [p["alpha"], p["beta"], p["gamma"], p["rho_hospital"], p["alpha2"], p["rho_imported"]]
    return np.asfortranarray(np.tile(np.array(row, dtype=np.float64), (n_sims, 1)))


def particle_parameters(n_particles: int, seed: int = 100) -> np.ndarray:
    """Parameter-inference sweep (BASELINE config 3): one (alpha, beta, gamma, rho_hosp, alpha2, rho_imp)
    particle per simulation, drawn from the priors of SURVEY.md section 8(d)."""
    rng = np.random.default_rng(seed)
    lo = np.array([0.05, 0.2, 0.02, 0.3, 0.1, 0.5])
    hi = np.array([0.30, 0.9, 0.20, 0.9, 0.5, 1.0])
    return np.asfortranarray(lo + (hi - lo) * rng.random((n_particles, 6)))
