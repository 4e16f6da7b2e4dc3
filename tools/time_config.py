"""Development: time the fused kernel on an arbitrary synthetic configuration.
usage: python tools/time_config.py wards census days members [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from abm_amro_b200 import engine, synth
nw, cen, nd, S = (int(x) for x in sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
wm, tp, n_init = synth.make_hospital(nw, cen, nd, seed=1)
par = synth.particle_parameters(S) if S > 4096 else synth.tutorial_parameters(S)
with engine.Topology(wm, tp, n_init) as topo:
    ms = []
    for _ in range(reps):
        r = topo.simulate(par, 0.2, 0.2, 2, 7, 1, 42)
        ms.append(r.kernel_ms)
    upd = wm.shape[0] * S
    print(f"{os.environ.get('AMRO_FAST_SHAPE','auto')} G={os.environ.get('AMRO_FAST_G','auto')}: {nw}w x {cen} x {nd}d x {S}: kernel {min(ms):.3f} ms, {upd / min(ms) / 1e6:.1f} G updates/s, frac {upd * (4 + 20 / S) / (min(ms) * 1e-3) / 6454.6e9:.3f}, total {r.total_ms:.2f} ms, checksum {int(r.counts_colonized.astype(np.int64).sum())}/{int(r.counts_detected.astype(np.int64).sum())}")
