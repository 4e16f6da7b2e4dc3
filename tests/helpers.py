"""Shared helpers of the parity tests (test infrastructure: may use oracle/)."""
import numpy as np

from oracle import oracle


def stream_to_tensors(icp, idp, wm, tp, par, ttd, tsh, tsa, stream):
    """Run the oracle on a flat uniform stream (the order the reference consumes it in) and record which uniform every
    (row, simulation, draw slot) consumed: the per-(row, sim) tensors the GPU replay mode takes.  Unused slots hold 2.0
    (a uniform that can never fire), never NaN, so a wrongly consumed slot shows up as a state mismatch."""
    R, S, n_init = wm.shape[0], icp.shape[1], icp.shape[0]
    rng = oracle.Rng(oracle.RNG_STREAM, stream=stream, record_shape=(R, S), record_init_shape=(n_init, S))
    out, p = oracle.simulate_discrete_model(icp, idp, wm, tp, par, ttd, tsh, tsa, 1, rng=rng, return_p=True)
    t = {k: np.where(np.isnan(v), 2.0, v) for k, v in
         dict(u_col=rng.rec_col, u_det=rng.rec_det, u_icol=rng.rec_icol, u_idet=rng.rec_idet).items()}
    return out, p, t, rng.stream_pos


def random_tensors(R, S, n_init, seed):
    g = np.random.default_rng(seed)
    return dict(u_col=g.random((R, S)), u_det=g.random((R, S)), u_icol=g.random((n_init, S)), u_idet=g.random((n_init, S)))


def expected_pressure(wm, n_init, state_cd, c0, topo_arrays):
    """Colonised count every ward-day starts with (the reference's sum(W) with unit weights), from a finished run:
    a record starts from its predecessor's post-step C (or its initial C on day 0, or 0)."""
    R = wm.shape[0]
    S = state_cd.shape[1] // 2
    C = state_cd[:, :S]
    pre = np.zeros((R, S))
    nd = wm[:, 5].astype(np.int64)
    src = np.nonzero(nd > 0)[0]
    pre[nd[src]] = C[src]
    init_rows = np.setdiff1d(np.arange(n_init), nd[src])
    pre[init_rows] = c0[init_rows]
    starts = topo_arrays["seg_row_start"]
    order = topo_arrays["order"]
    out = np.zeros((len(starts) - 1, S), dtype=np.int32)
    for s in range(len(starts) - 1):
        out[s] = pre[order[starts[s]:starts[s + 1]]].sum(0)
    return out
