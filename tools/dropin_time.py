"""Development: where the time of the drop-in simulate_and_collapse call goes on config 2 (run with AMRO_FAST_VERBOSE=1)."""
import sys, time, os
sys.path.insert(0, '.')
import numpy as np
from abm_amro_b200 import synth, amro
wm, tp, n_init = synth.make_hospital(30, 10000, 365, seed=1)
print("flags", wm.flags['F_CONTIGUOUS'], wm.dtype, tp.flags['F_CONTIGUOUS'])
for i in range(4):
    t=time.perf_counter()
    amro.simulate_and_collapse(wm, tp, 0.2, 0.2, n_init, .15, .55, .05, .6, .35, .9, 1000, 2, 7, 1, 42)
    print("call ms", (time.perf_counter()-t)*1e3, flush=True)
t=time.perf_counter(); m = wm[:,0].max(); print("np max ms", (time.perf_counter()-t)*1e3)
