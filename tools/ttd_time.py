"""Development: kernel time of config 2 for several time_to_detect values (the bit-sliced kernel's 3-bit countdown code from 3 on)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abm_amro_b200 import engine, synth
wm, tp, n_init = synth.make_hospital(30, 10_000, 365, seed=1)
par = synth.tutorial_parameters(1000)
with engine.Topology(wm, tp, n_init) as topo:
    for ttd in (2, 3, 6, 7):
        ms = min(topo.simulate(par, 0.2, 0.2, ttd, 7, 1, 42).kernel_ms for _ in range(4))
        r = topo.simulate(par, 0.2, 0.2, ttd, 7, 1, 42)
        print(f"time_to_detect {ttd}: kernel {ms:.3f} ms, kernel kind {r.kernel_kind}, detected on the last day {int(r.counts_detected[-1].sum())}")
