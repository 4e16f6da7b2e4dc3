"""Development probe: parity of both device paths vs the oracle on a few hospitals (run under gpurun)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from abm_amro_b200 import capi, engine, synth
from oracle import oracle

def case(nw, cen, nd, S, ps, ttd=2, tsh=7, tsa=1, seed=42, paths=(capi.PATH_FAST, capi.PATH_GENERAL)):
    wm, tp, n_init = synth.make_hospital(nw, cen, nd, seed=1, p_split=ps)
    par = synth.particle_parameters(S, seed=3)
    rng = np.random.default_rng(0)
    R = wm.shape[0]
    uc, ud = rng.random((R, S)), rng.random((R, S))
    uic, uid = rng.random((n_init, S)), rng.random((n_init, S))
    icp = np.full((n_init, S), 0.2); idp = np.full((n_init, S), 0.3)
    ref, refp = oracle.simulate_discrete_model(icp, idp, wm, tp, par, ttd, tsh, tsa, seed, return_p=True,
                                         rng=oracle.Rng(oracle.RNG_TENSOR, u_col=uc, u_det=ud, u_icol=uic, u_idet=uid))
    ref2 = oracle.simulate_discrete_model(icp, idp, wm, tp, par, ttd, tsh, tsa, seed, rng=oracle.Rng(oracle.RNG_PHILOX))
    cc, cd = oracle.total_positive_colonized(ref, S), oracle.total_positive_detected(ref, S)
    with engine.Topology(wm, tp, n_init, device=0) as topo:
        print(f"case wards={nw} census={cen} days={nd} S={S} split={ps}: R={R} fast_eligible={topo.info.fast_eligible} ({topo.info.why_not_fast.decode()})")
        for path in paths:
            if path == capi.PATH_FAST and not topo.info.fast_eligible:
                continue
            got = topo.simulate(par, icp, idp, ttd, tsh, tsa, seed, rng_mode=capi.RNG_REPLAY, path=path, want_state=True, want_p=True,
                                u_col=uc, u_det=ud, u_icol=uic, u_idet=uid)
            ok_state = np.array_equal(got.state, ref[:, 6:])
            ok_cnt = np.array_equal(got.counts_colonized, cc) and np.array_equal(got.counts_detected, cd)
            pm = ~np.isnan(refp)
            ok_p = np.array_equal(got.p[pm], refp[pm])
            free = topo.simulate(par, icp, idp, ttd, tsh, tsa, seed, path=path, want_state=True)
            ok_free = np.array_equal(free.state, ref2[:, 6:])
            ring = topo.simulate(par, icp, idp, ttd, tsh, tsa, seed, path=path)
            ok_ring = np.array_equal(ring.counts_colonized, oracle.total_positive_colonized(ref2, S)) and np.array_equal(ring.counts_detected, oracle.total_positive_detected(ref2, S))
            print(f"   path={got.path_used}: replay state {ok_state} counts {ok_cnt} P {ok_p} | philox state {ok_free} | summary-only counts {ok_ring} | kernel {got.kernel_ms:.3f} ms / {ring.kernel_ms:.3f} ms launches {got.launches}")
            if not ok_state:
                bad = np.argwhere(got.state != ref[:, 6:])
                print("      first mismatches", bad[:5], got.state[tuple(bad[0])], ref[:, 6:][tuple(bad[0])])

if __name__ == "__main__":
    case(1, 6, 3, 1, 0.0)
    case(5, 300, 30, 20, 0.0)
    case(3, 40, 12, 4, 0.3)
    case(30, 2000, 20, 64, 0.0)
    case(4, 100, 15, 9, 0.0, ttd=0, tsh=1, tsa=3)
