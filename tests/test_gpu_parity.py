"""GPU parity tests (run on the B200 box with `pytest -m gpu`).  Everything goes through the C ABI
(abm_amro_b200.capi / engine are ctypes over include/amro_b200.h; abm_amro_b200.amro is the pybind11 binding of it).

Gates (BASELINE.json north_star):
  1. replay: state bookkeeping, per-ward counts and per-day totals bit-exact against the reference (golden vectors
     generated from the compiled reference) when both consume the same pre-drawn uniforms;
  2. transition probabilities: max |P_gpu - P_ref| <= 1e-12 (they are in fact bit-identical);
  3. free running: bit-exact against the oracle's restatement of the Philox stream, and per-day KS tests at
     alpha = 0.01 against a 512-member ensemble of the reference (>= 95 % of (day, metric) tests must pass).
"""
import ctypes
import os

import numpy as np
import pytest

from abm_amro_b200 import capi, engine, ensemble, synth
from oracle import oracle
from tests import reference_cases
from tests.helpers import expected_pressure, random_tensors, stream_to_tensors

pytestmark = pytest.mark.gpu
P_TOL = 1e-12  # north_star: transition probabilities within 1e-12 in fp64


@pytest.fixture(params=["auto", "wide", "team"])
def shape(request, monkeypatch):
    """Launch shape of the fused kernel (kernels_fast.cu): chosen by the library, or forced."""
    if request.param != "auto":
        monkeypatch.setenv("AMRO_FAST_SHAPE", request.param[0])
    else:
        monkeypatch.delenv("AMRO_FAST_SHAPE", raising=False)
    return request.param


def _check_against(ref_out, ref_p, got, S):
    assert np.array_equal(got.state, ref_out[:, 6:]), "C / D columns differ"
    assert np.array_equal(got.counts_colonized, oracle.total_positive_colonized(ref_out, S))
    assert np.array_equal(got.counts_detected, oracle.total_positive_detected(ref_out, S))
    if got.p is not None:
        m = ~np.isnan(ref_p)
        assert np.max(np.abs(got.p[m] - ref_p[m]), initial=0.0) <= P_TOL
        assert np.array_equal(got.p[m], ref_p[m])  # the same IEEE operations in the same order: identical bits


# ---- 1 + 2: replay against the reference's golden outputs ----------------------------------------------------------
@pytest.mark.parametrize("name", ["replay_a", "replay_c"])
@pytest.mark.parametrize("path", [capi.PATH_FAST, capi.PATH_GENERAL])
def test_replay_matches_reference_golden(golden, name, path, shape):
    z = golden(name)
    ttd, tsh, tsa = (int(x) for x in z["args"])
    S = z["icp"].shape[1]
    out, p, t, used = stream_to_tensors(z["icp"], z["idp"], z["wm"], z["tp"], z["par"], ttd, tsh, tsa, z["stream"])
    assert np.array_equal(out, z["out"])                      # the oracle reproduces the reference's golden output
    with engine.Topology(z["wm"], z["tp"], z["icp"].shape[0]) as topo:
        assert topo.info.fast_eligible
        got = topo.simulate(z["par"], z["icp"], z["idp"], ttd, tsh, tsa, 1, rng_mode=capi.RNG_REPLAY, path=path,
                            want_state=True, want_p=True, **t)
        assert got.path_used == path
        _check_against(z["out"], p, got, S)                   # ... and so does the GPU, from the same uniforms


@pytest.mark.parametrize("path", [capi.PATH_FAST, capi.PATH_GENERAL])
def test_replay_half_weights_match_reference_golden(golden, path):
    # patients split over two wards weigh 1/2 in each: the fused kernel counts ward pressure in halves and keeps tables per
    # weight (a colonised half-weight record feels the ward's pressure too); the float64 path takes any weight
    z = golden("replay_b")
    ttd, tsh, tsa = (int(x) for x in z["args"])
    out, p, t, _ = stream_to_tensors(z["icp"], z["idp"], z["wm"], z["tp"], z["par"], ttd, tsh, tsa, z["stream"])
    assert set(np.unique(z["wm"][:, 4])) == {0.5, 1.0}
    with engine.Topology(z["wm"], z["tp"], z["icp"].shape[0]) as topo:
        assert topo.info.fast_eligible
        got = topo.simulate(z["par"], z["icp"], z["idp"], ttd, tsh, tsa, 1, rng_mode=capi.RNG_REPLAY, path=path, want_state=True, want_p=True, **t)
        assert got.path_used == path
        _check_against(z["out"], p, got, z["icp"].shape[1])


def test_other_fractional_weights_use_the_general_path():
    wm, tp, n_init = synth.make_hospital(3, 40, 8, seed=2, p_split=0.3)
    wm = wm.copy(order="F")
    wm[wm[:, 4] == 0.5, 4] = 0.25                                    # not 1 or 1/2: only the float64 path takes it
    S = 5
    par = synth.particle_parameters(S, seed=1)
    icp = np.full((n_init, S), 0.3); idp = np.full((n_init, S), 0.4)
    t = random_tensors(wm.shape[0], S, n_init, seed=3)
    ref, refp = oracle.simulate_discrete_model(icp, idp, wm, tp, par, 1, 2, 1, 5, return_p=True, rng=oracle.Rng(oracle.RNG_TENSOR, **t))
    with engine.Topology(wm, tp, n_init) as topo:
        assert not topo.info.fast_eligible and b"weights" in topo.info.why_not_fast
        got = topo.simulate(par, icp, idp, 1, 2, 1, 5, rng_mode=capi.RNG_REPLAY, want_state=True, want_p=True, **t)
        assert got.path_used == capi.PATH_GENERAL
        _check_against(ref, refp, got, S)
        with pytest.raises(ValueError, match="weights"):
            topo.simulate(par, icp, idp, 1, 2, 1, 5, path=capi.PATH_FAST)


CASES = [  # wards, census, days, S, ttd, tsh, tsa
    (1, 6, 3, 1, 0, 1, 1), (5, 300, 30, 20, 2, 7, 1), (30, 2000, 12, 64, 2, 7, 1), (4, 100, 15, 9, 0, 1, 3),
    (12, 150, 9, 33, 5, 2, 2), (2, 700, 6, 8, 1, 3, 1),
    (40, 100, 8, 17, 1, 2, 1),   # tiny wards: a 32-record trip spans more ward-days than a warp's table ring holds
    (5, 300, 30, 20, 6, 7, 1), (3, 200, 25, 40, 3, 2, 1), (4, 100, 15, 64, 4, 1, 3),   # time_to_detect 3 .. 6: the 3-bit countdown code of k_bits
    (6, 400, 12, 19, 2, 3, 1, 0.2), (3, 90, 10, 16, 0, 1, 1, 0.5), (4, 1200, 10, 70, 6, 7, 1, 0.3), (2, 50, 8, 33, 1, 2, 2, 0.6)]   # last entry: share of patients split over two wards (weight 1/2)


@pytest.mark.parametrize("case", CASES, ids=lambda c: "x".join(map(str, c)))
def test_replay_and_philox_match_oracle(case, shape):
    nw, cen, nd, S, ttd, tsh, tsa = case[:7]
    p_split = case[7] if len(case) > 7 else 0.0
    wm, tp, n_init = synth.make_hospital(nw, cen, nd, seed=nw + cen, p_split=p_split)
    par = synth.particle_parameters(S, seed=3)
    g = np.random.default_rng(1)
    icp, idp = g.random((n_init, S)) * 0.6, g.random((n_init, S))
    t = random_tensors(wm.shape[0], S, n_init, seed=2)
    ref, refp = oracle.simulate_discrete_model(icp, idp, wm, tp, par, ttd, tsh, tsa, 5, return_p=True,
                                               rng=oracle.Rng(oracle.RNG_TENSOR, **t))
    ref_free = oracle.simulate_discrete_model(icp, idp, wm, tp, par, ttd, tsh, tsa, 5, rng=oracle.Rng(oracle.RNG_PHILOX))
    with engine.Topology(wm, tp, n_init) as topo:
        for path in (capi.PATH_FAST, capi.PATH_GENERAL):
            got = topo.simulate(par, icp, idp, ttd, tsh, tsa, 5, rng_mode=capi.RNG_REPLAY, path=path, want_state=True,
                                want_p=True, want_pressure=(path == capi.PATH_FAST and p_split == 0.0), **t)
            _check_against(ref, refp, got, S)
            if path == capi.PATH_FAST and p_split == 0.0:   # per-ward colonisation-pressure counts (the reference's sum(W), :99)
                c0 = (t["u_icol"] < icp).astype(float)
                assert np.array_equal(got.pressure, expected_pressure(wm, n_init, got.state, c0, topo.debug_arrays()))
            free = topo.simulate(par, icp, idp, ttd, tsh, tsa, 5, path=path, want_state=True)
            assert np.array_equal(free.state, ref_free[:, 6:])
            assert np.array_equal(free.counts_colonized, oracle.total_positive_colonized(ref_free, S))
            summary_only = topo.simulate(par, icp, idp, ttd, tsh, tsa, 5, path=path, want_moments=True)  # two-day ring, no trajectory
            assert np.array_equal(summary_only.counts_colonized, free.counts_colonized)
            assert np.array_equal(summary_only.counts_detected, free.counts_detected)
            if path == capi.PATH_FAST and ttd <= 6:    # the bit-sliced kernel takes weights 1 and 1/2 and every time_to_detect up to 6
                assert summary_only.kernel_kind == capi.KERNEL_BITS
            c, d = free.counts_colonized.astype(float), free.counts_detected.astype(float)   # exact integers in float64
            assert np.array_equal(summary_only.moments, np.stack([c.sum(1), (c * c).sum(1), d.sum(1), (d * d).sum(1)], 1))


@pytest.mark.parametrize("g,barrier", [(3, "cluster"), (3, "global"), (8, "cluster"), (12, "global")])
def test_teams_of_several_ctas_match_the_oracle(monkeypatch, g, barrier):
    # small hospitals run one CTA per team; force teams of several CTAs (thread-block cluster barrier / counter in global
    # memory) and check the free-running stream and the replayed uniforms bit for bit, state included
    monkeypatch.setenv("AMRO_FAST_G", str(g))
    monkeypatch.setenv("AMRO_FAST_BARRIER", barrier)
    nw, cen, nd, S, ttd, tsh, tsa = 12, 3000, 10, 40, 2, 3, 1
    wm, tp, n_init = synth.make_hospital(nw, cen, nd, seed=21)
    par = synth.particle_parameters(S, seed=4)
    icp, idp = np.full((n_init, S), 0.25), np.full((n_init, S), 0.5)
    t = random_tensors(wm.shape[0], S, n_init, seed=6)
    ref = oracle.simulate_discrete_model(icp, idp, wm, tp, par, ttd, tsh, tsa, 3, rng=oracle.Rng(oracle.RNG_PHILOX))
    ref_replay = oracle.simulate_discrete_model(icp, idp, wm, tp, par, ttd, tsh, tsa, 3, rng=oracle.Rng(oracle.RNG_TENSOR, **t))
    with engine.Topology(wm, tp, n_init) as topo:
        free = topo.simulate(par, icp, idp, ttd, tsh, tsa, 3, path=capi.PATH_FAST, want_state=True)
        assert np.array_equal(free.state, ref[:, 6:])
        assert np.array_equal(free.counts_colonized, oracle.total_positive_colonized(ref, S))
        assert np.array_equal(free.counts_detected, oracle.total_positive_detected(ref, S))
        rep = topo.simulate(par, icp, idp, ttd, tsh, tsa, 3, rng_mode=capi.RNG_REPLAY, path=capi.PATH_FAST, want_state=True, **t)
        assert np.array_equal(rep.state, ref_replay[:, 6:])
        lean = topo.simulate(par, icp, idp, ttd, tsh, tsa, 3, path=capi.PATH_FAST)   # summary run: the lean instance of the same team
        assert np.array_equal(lean.counts_colonized, free.counts_colonized)
        assert np.array_equal(lean.counts_detected, free.counts_detected)


@pytest.mark.parametrize("path", [capi.PATH_FAST, capi.PATH_GENERAL])
def test_distance_to_data_is_computed_on_the_device(path):
    # a sweep hands back one number per particle -- the squared distance of its per-day counts to the data
    wm, tp, n_init = synth.make_hospital(4, 120, 20, seed=12)
    S = 37
    par = synth.particle_parameters(S, seed=8)
    g = np.random.default_rng(5)
    obs_c, obs_d = g.integers(0, 120, 20).astype(float), g.integers(0, 60, 20).astype(float)
    obs_c[3] = np.nan; obs_d[[0, 7]] = np.nan                     # days without data
    with engine.Topology(wm, tp, n_init) as topo:
        r = topo.simulate(par, 0.2, 0.3, 2, 7, 1, 9, path=path, observed_colonized=obs_c, observed_detected=obs_d)
        c, d = r.counts_colonized.astype(float), r.counts_detected.astype(float)
        want = np.nansum((c - obs_c[:, None]) ** 2, 0) + np.nansum((d - obs_d[:, None]) ** 2, 0)
        assert np.allclose(r.distance, want, rtol=1e-13, atol=0)  # sums of integers' squares: exact up to summation order
        only_d = topo.simulate(par, 0.2, 0.3, 2, 7, 1, 9, path=path, want_counts=False, observed_detected=obs_d)
        assert np.allclose(only_d.distance, np.nansum((d - obs_d[:, None]) ** 2, 0), rtol=1e-13, atol=0)


def test_degenerate_wards_and_extreme_parameters_match_the_oracle():
    # N = 0 gives an infinite or NaN force of infection and comparisons that are simply false; probabilities
    # above 1 and below 0 are not clamped; a schedule longer than the run never tests; ttd = 0 detects on the day of the test
    wm, tp, n_init = synth.make_hospital(6, 90, 9, seed=31, p_split=0.2)
    tp = tp.copy(order="F")
    tp[(tp[:, 1] == 2), 2] = 0.0                                      # ward 2: N = 0 every day -> F = inf (or NaN when nobody is colonised)
    S = 24
    par = synth.particle_parameters(S, seed=2)
    par[0] = [0.0, 5.0, 0.9, 1.0, 0.0, 1.0]                          # never clears, beta so large that P > 1
    par[1] = [1.0, 0.0, 0.0, 0.0, 1.0, 0.0]                          # always clears, nothing transmits, nothing is detected
    par[2] = [1.5, -0.3, -0.1, 2.0, -0.5, 0.5]                       # outside [0, 1] everywhere
    g = np.random.default_rng(3)
    icp, idp = g.random((n_init, S)), g.random((n_init, S))
    t = random_tensors(wm.shape[0], S, n_init, seed=9)
    for ttd, tsh, tsa in ((0, 1, 1), (3, 1000, 2), (7, 2, 1000)):
        with np.errstate(all="ignore"):
            ref, refp = oracle.simulate_discrete_model(icp, idp, wm, tp, par, ttd, tsh, tsa, 4, return_p=True,
                                                       rng=oracle.Rng(oracle.RNG_TENSOR, **t))
            ref_free = oracle.simulate_discrete_model(icp, idp, wm, tp, par, ttd, tsh, tsa, 4, rng=oracle.Rng(oracle.RNG_PHILOX))
        with engine.Topology(wm, tp, n_init) as topo:
            assert topo.info.fast_eligible
            for path in (capi.PATH_FAST, capi.PATH_GENERAL):
                got = topo.simulate(par, icp, idp, ttd, tsh, tsa, 4, rng_mode=capi.RNG_REPLAY, path=path, want_state=True, want_p=True, **t)
                assert np.array_equal(got.state, ref[:, 6:])
                assert np.array_equal(got.p, refp, equal_nan=True)      # inf and NaN probabilities included, bit for bit
                free = topo.simulate(par, icp, idp, ttd, tsh, tsa, 4, path=path, want_state=True)
                assert np.array_equal(free.state, ref_free[:, 6:])


def test_long_runs_flush_the_per_lane_counters(monkeypatch):
    # one CTA per team and 150,000 records a day: every warp steps more than 1024 trips a day, so the per-lane half-precision
    # day counters (exact up to 1024) are flushed in the middle of a run
    monkeypatch.setenv("AMRO_FAST_G", "1")
    wm, tp, n_init = synth.make_hospital(3, 150_000, 2, seed=4)
    S = 8
    par = synth.tutorial_parameters(S)
    ref = oracle.simulate_discrete_model(np.full((n_init, S), 0.3), np.full((n_init, S), 0.3), wm, tp, par, 2, 1, 1, 8,
                                         rng=oracle.Rng(oracle.RNG_PHILOX))
    with engine.Topology(wm, tp, n_init) as topo:
        got = topo.simulate(par, 0.3, 0.3, 2, 1, 1, 8, path=capi.PATH_FAST)
        assert np.array_equal(got.counts_colonized, oracle.total_positive_colonized(ref, S))
        assert np.array_equal(got.counts_detected, oracle.total_positive_detected(ref, S))


def test_unsorted_rows_and_sharded_ensembles_give_the_same_numbers():
    wm, tp, n_init = synth.make_hospital(4, 90, 10, seed=8)
    S = 24
    par = synth.particle_parameters(S, seed=4)
    icp, idp = np.full((n_init, S), 0.3), np.full((n_init, S), 0.4)
    with engine.Topology(wm, tp, n_init) as topo:
        whole = topo.simulate(par, icp, idp, 2, 7, 1, 11, want_state=True)
        # an ensemble sharded over "GPUs": global simulation index in the Philox counter -> identical columns
        for off in (0, 8, 16):
            part = topo.simulate(par[off:off + 8], icp[:, off:off + 8], idp[:, off:off + 8], 2, 7, 1, 11, sim_offset=off, want_state=True)
            assert np.array_equal(part.state[:, :8], whole.state[:, off:off + 8])
            assert np.array_equal(part.state[:, 8:], whole.state[:, S + off:S + off + 8])
            assert np.array_equal(part.counts_detected, whole.counts_detected[:, off:off + 8])
    # the same hospital with its rows (after the initial block) shuffled: records keep their identity
    g = np.random.default_rng(0)
    perm = np.r_[np.arange(n_init), n_init + g.permutation(wm.shape[0] - n_init)]
    inv = np.empty_like(perm); inv[perm] = np.arange(perm.size)
    sh = wm[perm].copy()
    m = sh[:, 5] > 0
    sh[m, 5] = inv[sh[m, 5].astype(int)]
    t = random_tensors(wm.shape[0], S, n_init, seed=6)
    ts = dict(t, u_col=t["u_col"][perm], u_det=t["u_det"][perm])
    with engine.Topology(wm, tp, n_init) as a, engine.Topology(sh, tp, n_init) as b:
        ra = a.simulate(par, icp, idp, 2, 7, 1, 11, rng_mode=capi.RNG_REPLAY, want_state=True, **t)
        rb = b.simulate(par, icp, idp, 2, 7, 1, 11, rng_mode=capi.RNG_REPLAY, want_state=True, **ts)
        assert not b.info.identity_order and rb.path_used == capi.PATH_FAST
        assert np.array_equal(rb.state, ra.state[perm]) and np.array_equal(rb.counts_colonized, ra.counts_colonized)
        ref = oracle.simulate_discrete_model(icp, idp, sh, tp, par, 2, 7, 1, 11, rng=oracle.Rng(oracle.RNG_PHILOX))
        assert np.array_equal(b.simulate(par, icp, idp, 2, 7, 1, 11, want_state=True).state, ref[:, 6:])


def test_exotic_links_fall_back_to_the_general_path_and_match_the_oracle():
    wm, tp, n_init = synth.make_hospital(3, 40, 8, seed=12)
    w2 = wm.copy()
    src = np.nonzero(w2[:, 5] > 0)[0]
    w2[src[5], 5] = src[5] + 1            # same-day target
    w2[src[40], 5] = 3                    # back to day 0
    w2[src[60], 5] = w2[src[61], 5]       # two sources, one target: the later row wins
    tp2 = tp[tp[:, 0] < 7]                # the last day is never stepped
    S = 5
    par = synth.particle_parameters(S, seed=1)
    icp, idp = np.full((n_init, S), 0.4), np.full((n_init, S), 0.5)
    t = random_tensors(wm.shape[0], S, n_init, seed=3)
    ref, refp = oracle.simulate_discrete_model(icp, idp, w2, tp2, par, 1, 2, 1, 3, return_p=True, rng=oracle.Rng(oracle.RNG_TENSOR, **t))
    with engine.Topology(w2, tp2, n_init) as topo:
        assert not topo.info.fast_eligible
        got = topo.simulate(par, icp, idp, 1, 2, 1, 3, rng_mode=capi.RNG_REPLAY, want_state=True, want_p=True, **t)
        _check_against(ref, refp, got, S)


# ---- the reference's own test-suite through the drop-in module -----------------------------------------------------------
@pytest.mark.parametrize("case", reference_cases.ALL, ids=lambda f: f.__name__)
def test_reference_suite_on_the_drop_in_module(case):
    from abm_amro_b200 import amro
    case(amro)


def test_single_step_entry_points_match_reference_golden(golden):
    z = golden("ward_step")
    N, S, ttd, dh, da = z["args"]
    S, n = int(S), z["wm"].shape[0]
    # per-(row, sim) uniforms from the flat stream: replay the oracle with recording
    rng = oracle.Rng(oracle.RNG_STREAM, stream=z["stream"], record_shape=(n, S))
    assert np.array_equal(oracle.progress_patients_probability_ward_1_timestep(z["wm"], float(N), z["par"], S, int(ttd), int(dh), int(da), rng=rng), z["out"])
    uc, ud = (np.asfortranarray(np.where(np.isnan(x), 2.0, x)) for x in (rng.rec_col, rng.rec_det))
    wmat, par = np.asfortranarray(z["wm"].copy()), np.asfortranarray(z["par"])
    capi.check(capi.lib().amro_progress_patients(wmat.ctypes.data, n, wmat.shape[1], n, None, 0, 0, 0, float(N), par.ctypes.data,
                                                 par.shape[0], 6, par.shape[0], S, int(ttd), int(dh), int(da), capi.RNG_REPLAY, 0, 0,
                                                 uc.ctypes.data, ud.ctypes.data, 0, None))
    assert np.array_equal(wmat, z["out"])                  # the caller's matrix is updated in place, like the reference
    z = golden("day_step")
    S, ttd, dh, da = (int(x) for x in z["args"])
    rng = oracle.Rng(oracle.RNG_STREAM, stream=z["stream"], record_shape=(n, S))
    oracle.progress_patients_1_timestep(z["wm"], z["tp"], z["par"], S, ttd, dh, da, rng=rng)
    uc, ud = (np.asfortranarray(np.where(np.isnan(x), 2.0, x)) for x in (rng.rec_col, rng.rec_det))
    wmat, tpd = np.asfortranarray(z["wm"].copy()), np.asfortranarray(z["tp"])
    capi.check(capi.lib().amro_progress_patients(wmat.ctypes.data, n, wmat.shape[1], n, tpd.ctypes.data, tpd.shape[0], 3, tpd.shape[0],
                                                 0.0, par.ctypes.data, par.shape[0], 6, par.shape[0], S, ttd, dh, da, capi.RNG_REPLAY,
                                                 0, 0, uc.ctypes.data, ud.ctypes.data, 0, None))
    assert np.array_equal(wmat, z["out"])
    from abm_amro_b200 import amro
    borrowed = np.asfortranarray(z["wm"].copy())
    ret = amro.progress_patients_1_timestep(borrowed, z["tp"], z["par"], S, ttd, dh, da)
    assert np.array_equal(ret, borrowed) and ret is not borrowed and not np.array_equal(borrowed, z["wm"])
    with pytest.raises(IndexError):
        amro.progress_patients_1_timestep(z["wm"].copy(order="F"), z["tp"][z["tp"][:, 1] != 2], z["par"], S, ttd, dh, da)


def test_totals_and_collapse_match_reference_golden(golden):
    from abm_amro_b200 import amro
    z = golden("summary")
    S = z["col"].shape[1]
    assert np.array_equal(amro.total_positive_colonized(z["full"], S), z["col"])
    assert np.array_equal(amro.total_positive_detected(z["full"], S), z["det"])
    n_init = int(z["n_init"])
    bug = amro.simulate_and_collapse(z["wm"], z["tp"], 0.2, 0.3, n_init, .15, .55, .05, .6, .35, .9, 40, 2, 7, 1, 5)
    fixed = amro.simulate_and_collapse(z["wm"], z["tp"], 0.2, 0.3, n_init, .15, .55, .05, .6, .35, .9, 40, 2, 7, 1, 5, bug_compatible=False)
    assert bug.shape == z["collapse"].shape == (30, 5) and np.all(bug[:, 3:] == 0) and np.array_equal(bug[:, 0], np.arange(30))
    assert np.array_equal(bug[:, 1:3], fixed[:, 3:5])                        # the reference's layout bug, reproduced
    ref = oracle.simulate_and_collapse(z["wm"], z["tp"], 0.2, 0.3, n_init, .15, .55, .05, .6, .35, .9, 40, 2, 7, 1, 5,
                                       rng=oracle.Rng(oracle.RNG_PHILOX))
    assert np.array_equal(bug, ref)                                           # same Philox stream, same arithmetic order


# ---- 3: free-running ensembles against the reference (statistical) ---------------------------------------------------------
def test_free_running_ensemble_matches_reference_distribution(golden):
    from scipy import stats
    z = golden("ensemble")
    nw, cen, nd, hseed = (int(x) for x in z["hospital"])
    ttd, tsh, tsa = (int(x) for x in z["args"])
    wm, tp, n_init = synth.make_hospital(nw, cen, nd, seed=hseed)
    S = z["col"].shape[1]
    with engine.Topology(wm, tp, n_init) as topo:
        got = topo.simulate(synth.tutorial_parameters(S), float(z["init"][0]), float(z["init"][1]), ttd, tsh, tsa, 424242)
    passed = total = 0
    for mine, ref in ((got.counts_colonized, z["col"]), (got.counts_detected, z["det"])):
        for d in range(nd):
            total += 1
            passed += stats.ks_2samp(mine[d], ref[d]).pvalue >= 0.01
        # means and variances within Monte Carlo error of two 512-member ensembles
        se = np.sqrt(mine.var(1) / S + ref.var(1) / S) + 1e-9
        assert np.all(np.abs(mine.mean(1) - ref.mean(1)) < 5 * se)
        assert np.all(np.abs(np.log((mine.var(1) + 1) / (ref.var(1) + 1))) < 0.6)
    assert passed / total >= 0.95, f"only {passed}/{total} per-day KS tests passed at alpha = 0.01"


# ---- full-size properties (BASELINE config 2: 30 wards x 10,000 patients x 365 days, 1,000 members) ---------------------
def test_full_size_config2_properties():
    wm, tp, n_init = synth.make_hospital(30, 10_000, 365, seed=1)
    S = 1000
    par = synth.tutorial_parameters(S)
    with engine.Topology(wm, tp, n_init) as topo:
        assert topo.info.fast_eligible and topo.info.n_rows == 3_650_000
        a = topo.simulate(par, 0.2, 0.2, 2, 7, 1, 42)
        b = topo.simulate(par, 0.2, 0.2, 2, 7, 1, 42)
        assert a.path_used == capi.PATH_FAST
        assert np.array_equal(a.counts_colonized, b.counts_colonized) and np.array_equal(a.counts_detected, b.counts_detected)
        assert np.all(a.counts_detected <= a.counts_colonized) and np.all(a.counts_colonized <= 10_000) and a.counts_colonized.min() > 0
        # a slice of the ensemble run on its own (another "GPU") reproduces its columns exactly
        part = topo.simulate(par[512:576], 0.2, 0.2, 2, 7, 1, 42, sim_offset=512)
        assert part.path_used == capi.PATH_FAST
        assert np.array_equal(part.counts_colonized, a.counts_colonized[:, 512:576])
        assert np.array_equal(part.counts_detected, a.counts_detected[:, 512:576])
        # another seed gives another ensemble with the same law: day-364 prevalence within 6 standard errors
        c = topo.simulate(par, 0.2, 0.2, 2, 7, 1, 43)
        assert not np.array_equal(a.counts_colonized, c.counts_colonized)
        x, y = a.counts_colonized[-1].astype(float), c.counts_colonized[-1].astype(float)
        assert abs(x.mean() - y.mean()) < 6 * np.sqrt(x.var() / S + y.var() / S)
        # the trajectory of 8 members, expanded to the reference's columns, agrees with the fused per-day totals
        traj = topo.simulate(par[:8], 0.2, 0.2, 2, 7, 1, 42, want_state=True)
        day = wm[:, 0].astype(int)
        col = np.stack([np.bincount(day, weights=(traj.state[:, s] == 1), minlength=365) for s in range(8)], 1)
        det = np.stack([np.bincount(day, weights=(traj.state[:, 8 + s] <= 0), minlength=365) for s in range(8)], 1)
        assert np.array_equal(col, a.counts_colonized[:, :8]) and np.array_equal(det, a.counts_detected[:, :8])


# ---- oracle parity at the benchmarked launch shapes ---------------------------------------------------------------------
# The throughput claims rest on particular launches: config 2 (16 teams of 18 CTAs on the global-memory barrier, the 3-CTAs-per-SM
# register budget), config 3 (one CTA per team, more teams than CTA slots: several waves) and a config-4-shaped network
# (4 teams of 148 CTAs).  Shards are bit-independent of each other (the Philox counters carry the GLOBAL simulation index),
# so two non-adjacent 32-member shards of the full run are compared bit for bit with the C oracle run on those members alone.
def _shard_counts_from_oracle(wm, tp, n_init, par, off, n, ttd, tsh, tsa, seed):
    ref = oracle.simulate_discrete_model(np.full((n_init, n), 0.2), np.full((n_init, n), 0.2), wm, tp, par[off:off + n], ttd, tsh, tsa, seed,
                                         rng=oracle.Rng(oracle.RNG_PHILOX, sim_offset=off))
    day, days = wm[:, 0].astype(int), int(wm[:, 0].max()) + 1
    col = np.stack([np.bincount(day, weights=(ref[:, 6 + s] == 1), minlength=days) for s in range(n)], 1)
    det = np.stack([np.bincount(day, weights=(ref[:, 6 + n + s] <= 0), minlength=days) for s in range(n)], 1)
    return col.astype(np.int32), det.astype(np.int32)


BENCH_SHAPES = {  # wards, census, days, members, particle sweep?, shard offsets, expected (CTAs per team, barrier, grid)
    "cfg2": (30, 10_000, 365, 1000, False, (32, 960), (18, capi.BARRIER_GLOBAL, 288)),
    "cfg3": (20, 5_000, 180, 65_536, True, (64, 40_000), (1, capi.BARRIER_CTA, 1024)),
    "cfg4_shaped": (500, 200_000, 10, 256, False, (0, 224), (148, capi.BARRIER_GLOBAL, 592)),
}


@pytest.mark.parametrize("name", list(BENCH_SHAPES))
def test_benchmarked_launch_shapes_match_the_oracle(name):
    nw, cen, nd, S, sweep, offsets, (g_want, barrier_want, grid_want) = BENCH_SHAPES[name]
    wm, tp, n_init = synth.make_hospital(nw, cen, nd, seed=1)
    par = synth.particle_parameters(S, seed=100) if sweep else synth.tutorial_parameters(S)
    with engine.Topology(wm, tp, n_init) as topo:
        got = topo.simulate(par, 0.2, 0.2, 2, 7, 1, 42)          # the default launch: what bench.py times
        assert got.path_used == capi.PATH_FAST and got.kernel_kind == capi.KERNEL_BITS
        assert (got.ctas_per_team, got.team_barrier, got.grid) == (g_want, barrier_want, grid_want)
        for off in offsets:
            col, det = _shard_counts_from_oracle(wm, tp, n_init, par, off, 32, 2, 7, 1, 42)
            assert np.array_equal(got.counts_colonized[:, off:off + 32], col), f"{name}: colonised counts of members {off}.. differ"
            assert np.array_equal(got.counts_detected[:, off:off + 32], det), f"{name}: detected counts of members {off}.. differ"


@pytest.mark.parametrize("members,instance", [(96, "roomy"), (30_016, "team")])
def test_large_wards_are_counted_in_several_flushes(members, instance):
    # The bit-sliced kernel PULLS a ward-day's pressure: it counts the colonised plane of the ward-day's records with a vertical
    # counter that holds 31, flushed as often as the ward is long.  Two wards of ~1500 records (47 chunks of 32) need several flushes
    # in both variants of the count: deep branch-free steps (the roomy instance: few CTAs) and lean groups (the 128-register
    # instance: more teams than three CTAs per SM).  Ward pressures and day totals against the oracle, bit for bit.
    wm, tp, n_init = synth.make_hospital(2, 3000, 6, seed=5)
    par = synth.particle_parameters(members, seed=3)
    with engine.Topology(wm, tp, n_init) as topo:
        got = topo.simulate(par, 0.2, 0.2, 2, 3, 1, 77, want_pressure=True)
        assert got.kernel_kind == capi.KERNEL_BITS
        assert (got.grid <= 3 * 148) == (instance == "roomy")
        for off in (0, members - 32):
            ref = oracle.simulate_discrete_model(np.full((n_init, 32), 0.2), np.full((n_init, 32), 0.2), wm, tp, par[off:off + 32], 2, 3, 1, 77,
                                                 rng=oracle.Rng(oracle.RNG_PHILOX, sim_offset=off))
            day = wm[:, 0].astype(int)
            col = np.stack([np.bincount(day, weights=(ref[:, 6 + s] == 1), minlength=6) for s in range(32)], 1)
            det = np.stack([np.bincount(day, weights=(ref[:, 38 + s] <= 0), minlength=6) for s in range(32)], 1)
            assert np.array_equal(got.counts_colonized[:, off:off + 32], col) and np.array_equal(got.counts_detected[:, off:off + 32], det)
        # the exported pressure = colonised patients every ward-day STARTS with: day 0 from the initial state, later days = what the
        # ward's records carried over; compared with the count of the oracle's previous-day states through the links
        nxt = wm[:, 5].astype(np.int64)
        seg = {}
        for i, (d, w) in enumerate(tp[:, :2]):
            seg[(int(d), int(w))] = i
        ref = oracle.simulate_discrete_model(np.full((n_init, 32), 0.2), np.full((n_init, 32), 0.2), wm, tp, par[:32], 2, 3, 1, 77,
                                             rng=oracle.Rng(oracle.RNG_PHILOX, sim_offset=0))
        want = np.zeros((tp.shape[0], 32), dtype=np.int64)
        src = np.nonzero(nxt >= 0)[0]
        for r in src:
            t = nxt[r]
            want[seg[(int(wm[t, 0]), int(wm[t, 1]))]] += (ref[r, 6:38] == 1)
        later = tp[:, 0] > 0
        assert np.array_equal(got.pressure[later, :32].astype(np.int64), want[later])


def test_ward_sized_ensemble_matches_reference_mean_and_variance(golden):
    # The daily draws of one ward-day and simulation share a threshold dither (rng.cuh): marginals are exact, and the joint law
    # is argued to be over-dispersed by at most n * 2^-14.  This checks it where it matters: wards of config 2's size (333
    # records per ward-day) against the compiled reference's own 512-member ensemble (tools/make_golden_wards.py), on the
    # colonised count every ward-day starts with -- the reference's sum(W) (:99), i.e. the per-ward counts.
    z = golden("ward_counts")
    nw, cen, nd, hseed = (int(x) for x in z["hospital"])
    ttd, tsh, tsa = (int(x) for x in z["args"])
    ref = z["counts"].astype(float)                                   # [ward-days of days >= 1 x 512]
    S = ref.shape[1]
    wm, tp, n_init = synth.make_hospital(nw, cen, nd, seed=hseed)
    with engine.Topology(wm, tp, n_init) as topo:
        got = topo.simulate(synth.tutorial_parameters(S), float(z["init"][0]), float(z["init"][1]), ttd, tsh, tsa, 99, want_pressure=True)
        assert got.kernel_kind == capi.KERNEL_BITS
    mine = got.pressure[int(z["first_segment"]):].astype(float)
    assert mine.shape == ref.shape
    se = np.sqrt(mine.var(1) / S + ref.var(1) / S)
    assert np.all(np.abs(mine.mean(1) - ref.mean(1)) < 5 * se)        # every ward-day's mean within Monte Carlo error
    lr = np.log(mine.var(1, ddof=1) / ref.var(1, ddof=1))             # log variance ratio: sd ~ sqrt(2/511 * 2) = 0.09 per ward-day
    assert np.all(np.abs(lr) < 0.5)
    assert abs(lr.mean()) < 5 * 0.0885 / np.sqrt(len(lr)), f"per-ward variances are off by {np.expm1(lr.mean()) * 100:.1f} % on average"


def test_compiled_hospital_cache_of_the_one_call_entry_points():
    # amro.simulate_and_collapse / simulate_discrete_model keep compiled hospitals, keyed by content
    from abm_amro_b200 import amro
    amro.clear_cache()
    base = capi.cache_stats()
    wm, tp, n_init = synth.make_hospital(4, 120, 12, seed=3)
    args = (0.2, 0.3, n_init, .15, .55, .05, .6, .35, .9, 24, 2, 7, 1, 5)
    a = amro.simulate_and_collapse(wm, tp, *args)
    s1 = capi.cache_stats()
    assert (s1["misses"] - base["misses"], s1["hits"] - base["hits"], s1["entries"]) == (1, 0, 1) and s1["bytes"] > 0
    b = amro.simulate_and_collapse(wm.copy(), tp.copy(), *args)          # other arrays, same content: a hit
    s2 = capi.cache_stats()
    assert (s2["misses"] - base["misses"], s2["hits"] - base["hits"]) == (1, 1) and np.array_equal(a, b)
    wm2 = wm.copy(order="F")
    wm2[wm2[:, 0] == 5, 3] = 1.0                                          # edited in place: every record of day 5 is an arrival
    c = amro.simulate_and_collapse(wm2, tp, *args)
    s3 = capi.cache_stats()
    assert (s3["misses"] - base["misses"], s3["entries"]) == (2, 2) and not np.array_equal(a, c)
    ref = oracle.simulate_and_collapse(wm2, tp, *args, rng=oracle.Rng(oracle.RNG_PHILOX))
    assert np.array_equal(c, ref)                                          # ... and the edit is what was simulated
    d = amro.simulate_and_collapse(wm, tp, 0.2, 0.3, n_init - 1, *args[3:])   # same tables, another initial block: another topology
    assert capi.cache_stats()["misses"] - base["misses"] == 3 and d.shape == a.shape
    for k in range(4):                                                     # more hospitals than slots: the oldest go
        w, t, n = synth.make_hospital(3, 50, 6, seed=20 + k)
        amro.simulate_and_collapse(w, t, 0.2, 0.3, n, *args[3:])
    s4 = capi.cache_stats()
    assert s4["entries"] == 4 and s4["misses"] - base["misses"] == 7
    amro.simulate_and_collapse(wm, tp, *args)                              # evicted meanwhile: compiled again, same numbers
    assert capi.cache_stats()["misses"] - base["misses"] == 8
    amro.clear_cache()
    assert capi.cache_stats()["entries"] == 0


def test_cache_speculation_never_trusts_the_quick_key():
    # A call whose quick key (pointers, shapes, a 1/256th sample of the contents) matches a cached hospital starts its run on that
    # hospital while the tables are hashed; the content hash decides.  One edited cell that the sample cannot see must still be
    # simulated: the speculative results are thrown away and the run is repeated on the newly compiled hospital.
    from abm_amro_b200 import amro
    amro.clear_cache()
    base = capi.cache_stats()
    wm, tp, n_init = synth.make_hospital(6, 400, 20, seed=8)   # 8000 rows x 6 columns: the sample reads 8 doubles out of every 187
    args = (0.2, 0.3, n_init, .15, .55, .05, .6, .35, .9, 32, 2, 7, 1, 9)
    a = amro.simulate_and_collapse(wm, tp, *args)
    b = amro.simulate_and_collapse(wm, tp, *args)               # same arrays again: speculative run, confirmed by the hash
    s = capi.cache_stats()
    assert (s["misses"] - base["misses"], s["hits"] - base["hits"]) == (1, 1) and np.array_equal(a, b)
    R = wm.shape[0]
    step = (R * 6) // 256
    flat = 3 * R + 5 * 400 + 100                                 # is_new of a day-5 record ...
    while flat % step < 8 or flat >= 6 * R - 8:                  # ... that is not among the sampled doubles
        flat += 1
    r, c = flat % R, flat // R
    assert c == 3
    wm[r, c] = 1.0 - wm[r, c]                                    # in place, same pointer, same shape, same sample
    e = amro.simulate_and_collapse(wm, tp, *args)
    s = capi.cache_stats()
    assert s["misses"] - base["misses"] == 2 and s["entries"] == 2
    ref = oracle.simulate_and_collapse(wm, tp, *args, rng=oracle.Rng(oracle.RNG_PHILOX))
    assert np.array_equal(e, ref) and not np.array_equal(e, a)
    amro.clear_cache()


# ---- hospitals generated on the device against the host compiler and the oracle ---------------------------
@pytest.mark.parametrize("shape", [(5, 300, 12), (40, 100, 8), (3, 2000, 5), (1, 33, 3)], ids=lambda s: "x".join(map(str, s)))
def test_device_generated_hospital_matches_host_compile_and_oracle(shape):
    spec = engine.hospital_spec(*shape, seed=7)
    wm, tp, n_init = engine.hospital_tables(spec)           # the same hospital as the reference's float64 tables
    S = 40
    par = synth.particle_parameters(S, seed=2)
    with engine.Topology.generate(spec) as g, engine.Topology(wm, tp, n_init) as h:
        assert h.info.fast_eligible and h.info.identity_order
        for f in ("n_rows", "n_rows_stepped", "n_init", "total_days", "n_days_out", "n_segments", "max_rows_per_day"):
            assert getattr(g.info, f) == getattr(h.info, f), f
        ga, ha = g.device_arrays(), h.device_arrays()        # what the kernels read: generated on the device == compiled by topology.cpp
        for k in ga:
            assert np.array_equal(ga[k], ha[k]), k
        a = g.simulate(par, 0.2, 0.3, 2, 7, 1, 5)
        b = h.simulate(par, 0.2, 0.3, 2, 7, 1, 5)
        assert a.kernel_kind == capi.KERNEL_BITS
        assert np.array_equal(a.counts_colonized, b.counts_colonized) and np.array_equal(a.counts_detected, b.counts_detected)
        ref = oracle.simulate_discrete_model(np.full((n_init, S), 0.2), np.full((n_init, S), 0.3), wm, tp, par, 2, 7, 1, 5,
                                             rng=oracle.Rng(oracle.RNG_PHILOX))
        assert np.array_equal(a.counts_colonized, oracle.total_positive_colonized(ref, S))
        assert np.array_equal(a.counts_detected, oracle.total_positive_detected(ref, S))
        traj = g.simulate(par, 0.2, 0.3, 5, 2, 1, 5, want_state=True)     # the trajectory kernel on a generated handle
        ref5 = oracle.simulate_discrete_model(np.full((n_init, S), 0.2), np.full((n_init, S), 0.3), wm, tp, par, 5, 2, 1, 5,
                                              rng=oracle.Rng(oracle.RNG_PHILOX))
        assert traj.kernel_kind == capi.KERNEL_CODES and np.array_equal(traj.state, ref5[:, 6:])
        with pytest.raises(ValueError, match="fused kernels only"):
            g.simulate(par, 0.2, 0.3, 2, 7, 1, 5, path=capi.PATH_GENERAL)


# ---- summary_of_total_positive on the device ------------------------------------------------------------
def test_device_summary_matches_the_host_routine_and_the_reference_golden(golden):
    z = golden("summary")
    n_days, S = z["col"].shape
    out = np.zeros((n_days, 3 + len(z["q"])), order="F")
    col = np.asfortranarray(z["col"])
    capi.check(capi.lib().amro_summary_of_total_positive_device(col.ctypes.data, n_days, S, n_days, z["q"].ctypes.data, len(z["q"]), 0, out.ctypes.data))
    assert np.array_equal(out, z["summ"])                               # the compiled reference's own output
    g = np.random.default_rng(5)
    qs = np.array([-0.1, 0.0, 1e-6, 0.025, 0.25, 0.5, 0.75, 0.975, 1.0 - 1e-6, 1.0, 1.2])
    for n_days, S, kind in ((7, 65_536, "counts"), (5, 1000, "reals"), (3, 2, "counts"), (4, 1, "reals"), (181, 4096, "counts")):
        x = g.integers(0, 5000, (n_days, S)).astype(float) if kind == "counts" else g.normal(size=(n_days, S)) * 1e3
        x = np.asfortranarray(x)
        host = np.zeros((n_days, 3 + len(qs)), order="F"); dev = np.zeros_like(host)
        capi.check(capi.lib().amro_summary_of_total_positive(x.ctypes.data, n_days, S, n_days, qs.ctypes.data, len(qs), host.ctypes.data))
        capi.check(capi.lib().amro_summary_of_total_positive_device(x.ctypes.data, n_days, S, n_days, qs.ctypes.data, len(qs), 0, dev.ctypes.data))
        assert np.array_equal(host, dev), (n_days, S, kind)             # the same IEEE operations in the same order
    from abm_amro_b200 import amro
    big = np.asfortranarray(g.integers(0, 9000, (180, 8192)).astype(float))     # large enough for the module to take the device path
    want = np.zeros((180, 3 + 3), order="F")
    capi.check(capi.lib().amro_summary_of_total_positive(big.ctypes.data, 180, 8192, 180, np.array([0.025, 0.5, 0.975]).ctypes.data, 3, want.ctypes.data))
    assert np.array_equal(amro.summary_of_total_positive(big, np.array([0.025, 0.5, 0.975])), want)
    # fused with a run: the summaries of the run's own counts never leave the device as counts
    wm, tp, n_init = synth.make_hospital(4, 200, 15, seed=6)
    par = synth.particle_parameters(300, seed=1)
    with engine.Topology(wm, tp, n_init) as topo:
        r = topo.simulate(par, 0.2, 0.3, 2, 7, 1, 9, quantiles=[0.05, 0.5, 0.95])
        assert np.array_equal(r.summary_colonized, oracle.summary_of_total_positive(r.counts_colonized.astype(float), np.array([0.05, 0.5, 0.95])))
        assert np.array_equal(r.summary_detected, oracle.summary_of_total_positive(r.counts_detected.astype(float), np.array([0.05, 0.5, 0.95])))


# ---- parameter-inference sweep ---------------------------------------------------------------------------
def test_abc_sweep_returns_one_distance_per_particle_and_finds_the_truth():
    from abm_amro_b200 import inference
    wm, tp, n_init = synth.make_hospital(6, 600, 40, seed=5)
    truth = np.array([[0.12, 0.75, 0.04, 0.7, 0.3, 0.8]])
    with engine.Topology(wm, tp, n_init) as topo:
        data = topo.simulate(np.repeat(truth, 32, 0), 0.2, 0.2, 2, 7, 1, 7)                 # "observed" series: the mean of 32 runs of the truth
        obs_c, obs_d = data.counts_colonized.mean(1), data.counts_detected.mean(1)
        obs_d[::9] = np.nan                                                                  # days without data
        n = 8192
        res = inference.abc_sweep(topo, obs_c, obs_d, n_particles=n, seed=11, keep=0.02)
        assert res.distance.shape == (n,) and res.particles.shape == (n, 6) and len(res.accepted) == round(0.02 * n)
        assert np.array_equal(res.particles, synth.particle_parameters(n, seed=100))         # the priors of SURVEY.md 8(d)
        # the distance of a block of particles, recomputed from their counts
        blk = slice(4096, 4096 + 64)
        r = topo.simulate(res.particles[blk], 0.2, 0.2, 2, 7, 1, 11, sim_offset=4096)
        c, d = r.counts_colonized.astype(float), r.counts_detected.astype(float)
        want = np.nansum((c - obs_c[:, None]) ** 2, 0) + np.nansum((d - obs_d[:, None]) ** 2, 0)
        assert np.allclose(res.distance[blk], want, rtol=1e-12, atol=0)
        # the accepted particles concentrate around the truth where the data are informative (beta: the force of infection)
        prior_err = abs(res.particles[:, 1].mean() - truth[0, 1])
        assert abs(res.posterior_mean[1] - truth[0, 1]) < 0.5 * prior_err
        assert res.posterior_sd[1] < 0.7 * res.particles[:, 1].std()
        assert np.all(np.diff(res.distance[res.accepted]) >= 0)


def test_device_pointer_api_matches_host_api():
    torch = pytest.importorskip("torch")
    wm, tp, n_init = synth.make_hospital(6, 400, 20, seed=2)
    S = 40
    par = synth.tutorial_parameters(S)
    dev = torch.device("cuda", 0)
    with engine.Topology(wm, tp, n_init) as topo:
        host = topo.simulate(par, 0.2, 0.2, 2, 7, 1, 9)
        par_d = torch.from_numpy(np.ascontiguousarray(par.T)).to(dev)
        icp_d = torch.tensor([0.2], dtype=torch.float64, device=dev)
        cc = torch.zeros((S, topo.info.n_days_out), dtype=torch.int32, device=dev); cd = torch.zeros_like(cc)
        topo.simulate_device(par_d.data_ptr(), S, S, icp_d.data_ptr(), icp_d.data_ptr(), cc.data_ptr(), cd.data_ptr(), S, 2, 7, 1, 9,
                             stream=torch.cuda.current_stream(dev).cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(cc.cpu().numpy().T, host.counts_colonized) and np.array_equal(cd.cpu().numpy().T, host.counts_detected)


def test_day_moments_peer_stores_match_the_all_reduce_arithmetic():
    # The run's collective as peer stores (amro_day_moments_peers): per-day sum / sum of squares of this GPU's counts, stored into
    # every destination.  On one GPU the destinations are three local buffers standing in for the peers' mapped memory; the
    # two-rank arithmetic (slots added in rank order) is what ensemble.PeerMoments.reduced returns on a real box.
    torch = pytest.importorskip("torch")
    import ctypes as C
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(3)
    n_days, S = 37, 203
    cc = torch.from_numpy(rng.integers(0, 900, size=(S, n_days)).astype(np.int32)).to(dev)      # == [n_days x S] column-major
    cd = torch.from_numpy(rng.integers(0, 300, size=(S, n_days)).astype(np.int32)).to(dev)
    bufs = [torch.full((4, n_days), -1.0, dtype=torch.float64, device=dev) for _ in range(3)]
    dst = (C.c_void_p * 3)(*[b.data_ptr() for b in bufs])
    capi.check(capi.lib().amro_day_moments_peers(cc.data_ptr(), cd.data_ptr(), n_days, S, n_days, dst, 3, 0,
                                                 torch.cuda.current_stream(dev).cuda_stream))
    torch.cuda.synchronize()
    want = ensemble.day_moments(cc, cd).T.contiguous()                    # [4 x n_days]: integers below 2^53, exact in any order
    for b in bufs:
        assert torch.equal(b, want)
    with pytest.raises(ValueError):
        capi.check(capi.lib().amro_day_moments_peers(cc.data_ptr(), cd.data_ptr(), n_days, S, n_days, dst, 17, 0, 0))


def test_half_weight_hospital_pressures_agree_between_the_fused_kernels(monkeypatch):
    # patients split over two wards weigh 1/2 in each: the bit-sliced kernel pulls sum(W) through two masks (full- and half-weight
    # records) and exports it in halves, like the 16-bit-code kernel's pushed counts; same draws, so the same numbers
    wm, tp, n_init = synth.make_hospital(5, 600, 12, seed=11, p_split=0.25)
    S = 70
    par = synth.particle_parameters(S, seed=2)
    with engine.Topology(wm, tp, n_init) as topo:
        assert topo.info.fast_eligible
        a = topo.simulate(par, 0.3, 0.4, 3, 2, 1, 17, want_pressure=True)
        assert a.kernel_kind == capi.KERNEL_BITS
        monkeypatch.setenv("AMRO_FAST_NOBITS", "1")
        b = topo.simulate(par, 0.3, 0.4, 3, 2, 1, 17, want_pressure=True)
        assert b.kernel_kind == capi.KERNEL_CODES
        assert np.array_equal(a.counts_colonized, b.counts_colonized) and np.array_equal(a.counts_detected, b.counts_detected)
        assert np.array_equal(a.pressure, b.pressure) and a.pressure.max() > 2     # in halves
