#!/usr/bin/env python
"""Generate tests/golden/ward_counts.npz from the UNMODIFIED reference (oracle/_ref, native RNG): the colonised count every
ward-day STARTS with (the reference's sum(W), abm_wards.cpp:99) for a 512-member ensemble on a hospital with config-2-sized
wards (6 wards x 2,000 patients: 333 records per ward-day, 30 days).  The GPU test compares the per-ward-day mean and variance
of its own free-running ensemble with these: the draws of a ward-day share a threshold dither (rng.cuh), and this is the
check that their joint law is the reference's at the ward sizes the benchmark runs.

Run in the build container (where /root/reference exists):  python tools/make_golden_wards.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from abm_amro_b200 import engine, synth  # noqa: E402
from oracle import ref_runner  # noqa: E402

HOSPITAL = (6, 2000, 30, 7)          # wards, census, days, generator seed
ARGS = (2, 7, 1)                      # time_to_detect, testing schedules
INIT = (0.2, 0.2)
S, CHUNKS = 64, 8

assert ref_runner.available("native"), "build oracle/_ref first (make -C oracle ref)"
wm, tp, n_init = synth.make_hospital(*HOSPITAL[:3], seed=HOSPITAL[3])
par = synth.tutorial_parameters(S)
icp = np.asfortranarray(np.full((n_init, S), INIT[0])); idp = np.asfortranarray(np.full((n_init, S), INIT[1]))
with engine.Topology(wm, tp, n_init, device=-1) as topo:   # host-only handle: the grouping of rows into ward-days
    arr = topo.debug_arrays()
starts, order, day_seg = arr["seg_row_start"], arr["order"], arr["day_seg_start"]
nd = wm[:, 5].astype(np.int64)
src = np.nonzero(nd > 0)[0]
first = int(day_seg[1])               # ward-days of day 0 start from the initial draw, which the output does not keep
chunks = []
for c in range(CHUNKS):
    out = ref_runner.run("simulate_discrete_model", [icp, idp, wm, tp, par, *ARGS, 2000 + c])["out"]
    C = out[:, 6:6 + S]
    pre = np.zeros_like(C)
    pre[nd[src]] = C[src]             # a record starts from its predecessor's post-step state (:402-407)
    counts = np.stack([pre[order[starts[s]:starts[s + 1]]].sum(0) for s in range(first, len(starts) - 1)])
    chunks.append(counts.astype(np.int16))
    print("chunk", c, counts.shape, counts.mean())
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ward_counts.npz"), hospital=np.array(HOSPITAL), args=np.array(ARGS),
                    init=np.array(INIT), first_segment=np.array(first), counts=np.concatenate(chunks, axis=1))
print("wrote ward_counts", np.concatenate(chunks, axis=1).shape)
