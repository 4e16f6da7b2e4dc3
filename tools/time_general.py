"""Development: the float64 (GENERAL) path on a hospital with half-weight records, forced with path = PATH_GENERAL."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from abm_amro_b200 import capi, engine, synth
for (nw, cen, nd, S, ps) in [(30, 10000, 60, 256, 0.05), (30, 10000, 60, 1000, 0.05)]:
    wm, tp, n_init = synth.make_hospital(nw, cen, nd, seed=1, p_split=ps)
    par = synth.tutorial_parameters(S)
    with engine.Topology(wm, tp, n_init) as topo:
        print("fast eligible:", topo.info.fast_eligible, topo.info.why_not_fast)
        for _ in range(2):
            r = topo.simulate(par, 0.2, 0.2, 2, 7, 1, 42, path=capi.PATH_GENERAL)
        upd = wm.shape[0] * S
        f = topo.simulate(par, 0.2, 0.2, 2, 7, 1, 42)
        f = topo.simulate(par, 0.2, 0.2, 2, 7, 1, 42)
        print(f"fused kernel kind {f.kernel_kind} (2 = bit-sliced): kernel {f.kernel_ms:.2f} ms; counts equal to the float64 path: {bool((f.counts_colonized == r.counts_colonized).all() and (f.counts_detected == r.counts_detected).all())}")
        print(f"{("fast" if r.path_used == capi.PATH_FAST else "general")} path {nw}w x {cen} x {nd}d x {S}: kernel {r.kernel_ms:.2f} ms, {upd / r.kernel_ms / 1e6:.1f} G updates/s, launches {r.launches}")
