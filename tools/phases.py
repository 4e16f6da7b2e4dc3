"""Development: per-phase cycle breakdown of the fast kernel (needs a variant built with -DAMRO_PROFILE_PHASES;
AMRO_B200_LIB selects it, default lib_variants/phases)."""
import ctypes, os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
os.environ.setdefault("AMRO_B200_LIB", os.path.join(root, "abm_amro_b200/lib_variants/phases/libamro_b200.so"))
import numpy as np
from abm_amro_b200 import capi, engine, synth
wm, tp, n_init = synth.make_hospital(30, 10000, 365, seed=1)
S = 1000
par = synth.tutorial_parameters(S)
topo = engine.Topology(wm, tp, n_init, device=0)
L = capi.lib()
buf = (ctypes.c_ulonglong * 8)()
for it in range(3):
    r = topo.simulate(par, 0.2, 0.2, 2, 7, 1, 42)
    L.amro_debug_phase_cycles(buf, 1)
    v = np.array(list(buf), dtype=np.float64)
    n_cta = int(os.environ.get("N_CTA", "125"))
    print(f"kernel {r.kernel_ms:.3f} ms; per-CTA cycles (thread 0): table {v[0]/n_cta:.0f}  records {v[1]/n_cta:.0f}  flush {v[2]/n_cta:.0f}  wait {v[3]/n_cta:.0f}  arrive {v[4]/n_cta:.0f}  setup {v[5]/n_cta:.0f}  total {v[:6].sum()/n_cta:.0f} (= {v[:6].sum()/n_cta/1.965e6:.2f} ms)")
