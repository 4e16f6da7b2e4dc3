"""Development: what a drop-in user sees -- repeated amro.simulate_and_collapse calls on one hospital."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from abm_amro_b200 import amro, synth
wm, tp, n_init = synth.make_hospital(30, 10000, 365, seed=1)
for i in range(5):
    t0 = time.perf_counter()
    out = amro.simulate_and_collapse(wm, tp, 0.2, 0.2, n_init, .15 + 0.01 * i, .55, .05, .6, .35, .9, 1000, 2, 7, 1, 42)
    print(f"call {i}: {(time.perf_counter() - t0) * 1e3:.1f} ms  -> {3.65e9 / (time.perf_counter() - t0) / 1e9:.1f} G updates/s", out[180, 1:3])
amro.clear_cache()
