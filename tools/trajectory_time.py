"""Development / profiles: time of the drop-in amro.simulate_discrete_model on a 64-member slice of config 2 (the full
[R x (6 + 2S)] float64 matrix of the reference: 3.9 GB) and the rate at which the trajectory leaves the device."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from abm_amro_b200 import amro, engine, synth

S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
wm, tp, n_init = synth.make_hospital(30, 10_000, 365, seed=1)
par = synth.tutorial_parameters(S)
icp = np.full((n_init, S), 0.2, order="F"); idp = np.full((n_init, S), 0.2, order="F")
for rep in range(3):
    t = time.perf_counter()
    out = amro.simulate_discrete_model(icp, idp, wm, tp, par, 2, 7, 1, 42)
    dt = time.perf_counter() - t
    gb = out.nbytes / 1e9
    print(f"amro.simulate_discrete_model: {out.shape} float64 = {gb:.2f} GB in {dt * 1e3:.0f} ms = {gb / dt:.2f} GB/s", flush=True)
    del out
with engine.Topology(wm, tp, n_init) as topo:
    for rep in range(3):
        t = time.perf_counter()
        r = topo.simulate(par, 0.2, 0.2, 2, 7, 1, 42, want_state=True, want_counts=False)
        dt = time.perf_counter() - t
        gb = r.state.nbytes / 1e9
        print(f"handle API (hospital already compiled): {gb:.2f} GB of C / D columns in {dt * 1e3:.0f} ms = {gb / dt:.2f} GB/s (kernel {r.kernel_ms:.1f} ms)", flush=True)
