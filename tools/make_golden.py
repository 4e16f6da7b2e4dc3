#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference compiled under oracle/_ref (oracle/Makefile `ref`).

Run in the build container (where /root/reference exists); the fixtures are committed so that the GPU box, which has
no /root/reference, can still pin the oracle and the CUDA path against the reference's own outputs.

  native_*   simulate_discrete_model with the reference's native RNG (glibc rand + mt19937_64 fill), given seed
  replay_*   the ARMA_RNG_ALT build fed a flat uniform stream: inputs, stream, output, number of uniforms consumed
  ward_*     progress_patients_probability_ward_1_timestep / progress_patients_1_timestep on replayed streams
  ensemble   per-day colonised / detected counts of a 512-member reference ensemble (for the KS test of the
             free-running GPU ensemble)
  summary    total_positive_* / summary_of_total_positive / simulate_and_collapse outputs
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from abm_amro_b200 import synth  # noqa: E402
from oracle import ref_runner  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
assert ref_runner.available("native") and ref_runner.available("replay"), "build oracle/_ref first (make -C oracle ref)"


def save(name, **kw):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **kw)
    print("wrote", name, {k: getattr(v, "shape", v) for k, v in kw.items()})


# ---- native-RNG runs -------------------------------------------------------------------------------------------
for tag, (nw, cen, nd, S, ps, ttd, tsh, tsa, seed) in {
    "native_a": (5, 60, 30, 6, 0.0, 2, 7, 1, 42),
    "native_b": (3, 40, 12, 4, 0.3, 0, 1, 3, 7),       # fractional weights, ttd = 0
    "native_c": (2, 30, 8, 1, 0.0, 3, 2, 2, 34279),    # single simulation
}.items():
    wm, tp, n_init = synth.make_hospital(nw, cen, nd, seed=1, p_split=ps)
    par = synth.particle_parameters(S, seed=5)
    icp = np.asfortranarray(np.full((n_init, S), 0.2)); idp = np.asfortranarray(np.full((n_init, S), 0.3))
    out = ref_runner.run("simulate_discrete_model", [icp, idp, wm, tp, par, ttd, tsh, tsa, seed])["out"]
    save(tag, wm=wm, tp=tp, par=par, icp=icp, idp=idp, args=np.array([ttd, tsh, tsa, seed]), out=out)

# ---- replayed-stream runs ------------------------------------------------------------------------------------------
for tag, (nw, cen, nd, S, ps, ttd, tsh, tsa) in {
    "replay_a": (5, 60, 30, 6, 0.0, 2, 7, 1),
    "replay_b": (3, 40, 12, 4, 0.3, 1, 2, 1),          # fractional weights
    "replay_c": (4, 50, 10, 9, 0.0, 0, 1, 3),
}.items():
    wm, tp, n_init = synth.make_hospital(nw, cen, nd, seed=2, p_split=ps)
    par = synth.particle_parameters(S, seed=9)
    rng = np.random.default_rng(11)
    icp = np.asfortranarray(rng.random((n_init, S)) * 0.5); idp = np.asfortranarray(rng.random((n_init, S)))
    stream = rng.random(2 * n_init * S + 2 * wm.shape[0] * S + 16)
    r = ref_runner.run("simulate_discrete_model", [icp, idp, wm, tp, par, ttd, tsh, tsa, 1], variant="replay", stream=stream)
    used = int(r["stream_pos"])
    save(tag, wm=wm, tp=tp, par=par, icp=icp, idp=idp, args=np.array([ttd, tsh, tsa]), stream=stream[:used], out=r["out"], used=np.array(used))

# ---- single ward / single day steps on arbitrary float64 state -------------------------------------------------------
rng = np.random.default_rng(3)
n, S = 40, 5
wmat = np.zeros((n, 6 + 2 * S), order="F")
wmat[:, 1] = rng.integers(1, 4, n); wmat[:, 2] = np.arange(n); wmat[:, 3] = rng.integers(0, 2, n)
wmat[:, 4] = rng.choice([1.0, 0.5, 0.25], n); wmat[:, 5] = -1
wmat[:, 6:6 + S] = rng.integers(0, 2, (n, S))
wmat[:, 6 + S:] = rng.choice([np.inf, 4, 2, 1, 0, -1, -3], (n, S))
par = synth.particle_parameters(S, seed=2)
tpd = np.asfortranarray(np.array([[0, 1, 13.0], [0, 2, 15.5], [0, 3, 11.0], [0, 1, 99.0]]))
stream = rng.random(2 * n * S)
r = ref_runner.run("progress_patients_probability_ward_1_timestep", [wmat.copy(order="F"), 37.5, par, S, 2, 1, 0], variant="replay", stream=stream)
save("ward_step", wm=wmat, par=par, args=np.array([37.5, S, 2, 1, 0]), stream=stream[:int(r["stream_pos"])], out=r["out"], mutated=r["mutated_input"])
r = ref_runner.run("progress_patients_1_timestep", [wmat.copy(order="F"), tpd, par, S, 1, 1, 1], variant="replay", stream=stream)
save("day_step", wm=wmat, tp=tpd, par=par, args=np.array([S, 1, 1, 1]), stream=stream[:int(r["stream_pos"])], out=r["out"], mutated=r["mutated_input"])

# ---- summaries ------------------------------------------------------------------------------------------------------
z = np.load(os.path.join(OUT, "native_a.npz"))
full, S = z["out"], z["par"].shape[0]
col = ref_runner.run("total_positive_colonized", [full, S])["out"]
det = ref_runner.run("total_positive_detected", [full, S])["out"]
summ = ref_runner.run("summary_of_total_positive", [col, np.array([0.025, 0.2, 0.5, 0.9, 0.975])])["out"]
wm, tp, n_init = synth.make_hospital(5, 60, 30, seed=1)
coll = ref_runner.run("simulate_and_collapse", [wm, tp, 0.2, 0.3, n_init, .15, .55, .05, .6, .35, .9, 40, 2, 7, 1, 2384235])["out"]
save("summary", full=full, col=col, det=det, q=np.array([0.025, 0.2, 0.5, 0.9, 0.975]), summ=summ,
     wm=wm, tp=tp, n_init=np.array(n_init), collapse=coll)

# ---- reference ensemble for the statistical (KS) parity test ----------------------------------------------------------
wm, tp, n_init = synth.make_hospital(6, 240, 60, seed=4)
S, chunks = 64, 8
par = synth.tutorial_parameters(S)
icp = np.asfortranarray(np.full((n_init, S), 0.2)); idp = np.asfortranarray(np.full((n_init, S), 0.2))
cols, dets = [], []
for c in range(chunks):
    out = ref_runner.run("simulate_discrete_model", [icp, idp, wm, tp, par, 2, 7, 1, 1000 + c])["out"]
    cols.append(ref_runner.run("total_positive_colonized", [out, S])["out"])
    dets.append(ref_runner.run("total_positive_detected", [out, S])["out"])
save("ensemble", hospital=np.array([6, 240, 60, 4]), args=np.array([2, 7, 1]), init=np.array([0.2, 0.2]),
     col=np.concatenate(cols, axis=1).astype(np.int32), det=np.concatenate(dets, axis=1).astype(np.int32))
