"""Development: the regional / national configurations of BASELINE.json at full census, host-generated.
usage: python tools/time_scale.py wards census days members [reps]     (config 4: 500 1000000 365 256)
Prints the generation / topology-compile / kernel times and the roofline fraction at 4 + 20/S bytes per update."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from abm_amro_b200 import engine, synth
nw, cen, nd, S = (int(x) for x in sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
t0 = time.time()
wm, tp, n_init = synth.make_hospital(nw, cen, nd, seed=1)
t1 = time.time()
print(f"generated {wm.shape[0]} patient-day rows ({wm.nbytes / 1e9:.1f} GB) in {t1 - t0:.1f} s", flush=True)
par = synth.tutorial_parameters(S)
with engine.Topology(wm, tp, n_init) as topo:
    t2 = time.time()
    print(f"topology compiled and uploaded in {t2 - t1:.1f} s; fast path: {bool(topo.info.fast_eligible)}", flush=True)
    ms = []
    for _ in range(reps):
        r = topo.simulate(par, 0.2, 0.2, 2, 7, 1, 42)
        ms.append(r.kernel_ms)
    upd = wm.shape[0] * S
    print(f"{nw}w x {cen} x {nd}d x {S}: kernel {min(ms):.3f} ms, {upd / min(ms) / 1e6:.1f} G updates/s, frac {upd * (4 + 20 / S) / (min(ms) * 1e-3) / 6454.6e9:.3f}, "
          f"path {r.path_used}, prevalence last day {r.counts_colonized[-1].mean() / cen:.4f}, checksum {int(r.counts_colonized.astype(np.int64).sum())}/{int(r.counts_detected.astype(np.int64).sum())}")
