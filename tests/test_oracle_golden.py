"""The CPU oracle pinned against the reference: golden vectors generated from the compiled, unmodified reference
(tools/make_golden.py), the reference's own 14 test cases, and -- where oracle/_ref is present -- the reference itself."""
import numpy as np
import pytest

from abm_amro_b200 import synth
from oracle import oracle, ref_runner
from tests.helpers import stream_to_tensors


@pytest.mark.parametrize("name", ["native_a", "native_b", "native_c"])
def test_native_rng_runs_match_reference(golden, name):
    z = golden(name)
    ttd, tsh, tsa, seed = (int(x) for x in z["args"])
    out = oracle.simulate_discrete_model(z["icp"], z["idp"], z["wm"], z["tp"], z["par"], ttd, tsh, tsa, seed)
    assert np.array_equal(out, z["out"])


@pytest.mark.parametrize("name", ["replay_a", "replay_b", "replay_c"])
def test_replayed_stream_runs_match_reference(golden, name):
    z = golden(name)
    ttd, tsh, tsa = (int(x) for x in z["args"])
    rng = oracle.Rng(oracle.RNG_STREAM, stream=z["stream"])
    out = oracle.simulate_discrete_model(z["icp"], z["idp"], z["wm"], z["tp"], z["par"], ttd, tsh, tsa, 1, rng=rng)
    assert np.array_equal(out, z["out"])
    assert rng.stream_pos == int(z["used"]) == z["stream"].size  # same number of uniforms, in the same order


@pytest.mark.parametrize("name", ["replay_a", "replay_b", "replay_c"])
def test_tensor_replay_equals_stream_replay(golden, name):
    z = golden(name)
    ttd, tsh, tsa = (int(x) for x in z["args"])
    out, p, t, used = stream_to_tensors(z["icp"], z["idp"], z["wm"], z["tp"], z["par"], ttd, tsh, tsa, z["stream"])
    rng = oracle.Rng(oracle.RNG_TENSOR, **t)
    out2, p2 = oracle.simulate_discrete_model(z["icp"], z["idp"], z["wm"], z["tp"], z["par"], ttd, tsh, tsa, 1, rng=rng, return_p=True)
    assert np.array_equal(out2, z["out"])
    assert np.array_equal(p2, p)


def test_ward_and_day_step_match_reference(golden):
    z = golden("ward_step")
    N, S, ttd, dh, da = z["args"]
    got = oracle.progress_patients_probability_ward_1_timestep(z["wm"], float(N), z["par"], int(S), int(ttd), int(dh), int(da),
                                                               rng=oracle.Rng(oracle.RNG_STREAM, stream=z["stream"]))
    assert np.array_equal(got, z["out"]) and np.array_equal(got, z["mutated"])
    z = golden("day_step")
    S, ttd, dh, da = (int(x) for x in z["args"])
    got = oracle.progress_patients_1_timestep(z["wm"], z["tp"], z["par"], S, ttd, dh, da,
                                              rng=oracle.Rng(oracle.RNG_STREAM, stream=z["stream"]))
    assert np.array_equal(got, z["out"])


def test_summaries_match_reference(golden):
    z = golden("summary")
    S = z["col"].shape[1]
    assert np.array_equal(oracle.total_positive_colonized(z["full"], S), z["col"])
    assert np.array_equal(oracle.total_positive_detected(z["full"], S), z["det"])
    assert np.array_equal(oracle.summary_of_total_positive(z["col"], z["q"]), z["summ"])
    got = oracle.simulate_and_collapse(z["wm"], z["tp"], 0.2, 0.3, int(z["n_init"]), .15, .55, .05, .6, .35, .9, 40, 2, 7, 1, 2384235)
    assert np.array_equal(got, z["collapse"])
    assert np.all(got[:, 3:] == 0)  # the reference's column bug: sd lands in columns 1-2, 3-4 stay zero


def test_philox_known_answers():
    # Random123 known-answer vectors for philox4x32-10
    assert oracle.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_threshold_map():
    assert oracle.threshold(0.0) == 0 and oracle.threshold(-1.0) == 0 and oracle.threshold(float("nan")) == 0
    assert oracle.threshold(1.0) == 2 ** 32 and oracle.threshold(7.5) == 2 ** 32
    assert oracle.threshold(0.5) == 2 ** 31 and oracle.threshold(2.0 ** -32) == 1
    for p in np.random.default_rng(0).random(200):
        t = oracle.threshold(p)
        assert (t - 1) * 2.0 ** -32 < p <= t * 2.0 ** -32  # number of r with r * 2^-32 < p


def test_dithered_daily_threshold_is_exact_in_expectation():
    # RNG spec of the daily step (abm_amro_b200/csrc/rng.cuh): fires iff v14 < B(p), B(p) = (T30(p) + r16) >> 16.
    # Averaged over the 2^16 dithers the firing probability is T30(p) / 2^30 exactly; "never" and "always" have no dither.
    assert oracle.threshold30(0.0) == 0 and oracle.threshold30(-3.0) == 0 and oracle.threshold30(float("nan")) == 0
    assert oracle.threshold30(1.0) == 2 ** 30 and oracle.threshold30(1.5) == 2 ** 30 and oracle.threshold30(float("inf")) == 2 ** 30
    assert oracle.threshold30(0.5) == 2 ** 29 and oracle.threshold30(2.0 ** -30) == 1 and oracle.threshold30(2.0 ** -31) == 1
    for r in (0, 1, 12345, 65535):
        assert oracle.bernoulli14(0.0, r) == 0 and oracle.bernoulli14(1.0, r) == 2 ** 14
    rs = np.arange(65536, dtype=np.int64)
    for p in list(np.random.default_rng(1).random(20)) + [1e-9, 0.25, 0.999999999, 1 - 2.0 ** -30]:
        t = oracle.threshold30(p)
        assert (t - 1) * 2.0 ** -30 < p <= t * 2.0 ** -30
        b = (t + rs) >> 16                                   # what the oracle / the kernels compute per dither
        assert b.min() >= 0 and b.max() <= 2 ** 14
        assert int(b.sum()) == t                             # sum_r P(v14 < B) * 2^14 = T30: exact marginal
        for r in (0, 777, 65535):
            assert oracle.bernoulli14(p, r) == int((t + r) >> 16)


def test_philox_chunks_are_the_halves_of_the_words():
    seed, row, stream = 0x1234567890ABCDEF, 987654321, 6
    for group in (0, 5):
        words = oracle.philox4x32_10([row & 0xFFFFFFFF, row >> 32, group, stream], [seed & 0xFFFFFFFF, seed >> 32])
        for k in range(8):
            assert oracle.philox_chunk(seed, row, group * 8 + k, stream) == (words[k >> 1] >> (16 * (k & 1))) & 0xFFFF


def test_philox_mode_dither_follows_the_ward_day():
    # the dither index is the ward-day in the reference's visiting order (day ascending, ward ascending), so a run is
    # reproduced by stepping its days one at a time only if the ward-day counter carries on
    wm, tp, n_init = synth.make_hospital(3, 30, 4, seed=9)
    par = synth.particle_parameters(8, seed=2)
    icp = np.full((n_init, 8), 0.3); idp = np.full((n_init, 8), 0.4)
    a = oracle.simulate_discrete_model(icp, idp, wm, tp, par, 1, 2, 1, 5, rng=oracle.Rng(oracle.RNG_PHILOX))
    b = oracle.simulate_discrete_model(icp, idp, wm, tp, par, 1, 2, 1, 5, rng=oracle.Rng(oracle.RNG_PHILOX))
    assert np.array_equal(a, b)
    c = oracle.simulate_discrete_model(icp, idp, wm, tp, par, 1, 2, 1, 6, rng=oracle.Rng(oracle.RNG_PHILOX))
    assert not np.array_equal(a[:, 6:], c[:, 6:])


def test_philox_mode_is_a_function_of_row_and_global_simulation():
    wm, tp, n_init = synth.make_hospital(3, 40, 8, seed=3)
    par = synth.particle_parameters(12, seed=1)
    icp = np.full((n_init, 12), 0.3); idp = np.full((n_init, 12), 0.4)
    full = oracle.simulate_discrete_model(icp, idp, wm, tp, par, 1, 2, 1, 99, rng=oracle.Rng(oracle.RNG_PHILOX))
    part = oracle.simulate_discrete_model(icp[:, 8:], idp[:, 8:], wm, tp, par[8:], 1, 2, 1, 99,
                                          rng=oracle.Rng(oracle.RNG_PHILOX, sim_offset=8))
    assert np.array_equal(part[:, 6:10], full[:, 6 + 8:6 + 12]) and np.array_equal(part[:, 10:], full[:, 18 + 8:])


@pytest.mark.skipif(not ref_runner.available("native"), reason="oracle/_ref not built (no /root/reference here)")
def test_live_reference_on_fresh_hospitals():
    for nw, cen, nd, S, ps, seed in [(4, 80, 20, 5, 0.0, 5), (3, 50, 10, 3, 0.25, 6)]:
        wm, tp, n_init = synth.make_hospital(nw, cen, nd, seed=seed, p_split=ps)
        par = synth.particle_parameters(S, seed=seed)
        icp = np.asfortranarray(np.full((n_init, S), 0.25)); idp = np.asfortranarray(np.full((n_init, S), 0.5))
        ref = ref_runner.run("simulate_discrete_model", [icp, idp, wm, tp, par, 2, 3, 1, 77 + seed])["out"]
        assert np.array_equal(oracle.simulate_discrete_model(icp, idp, wm, tp, par, 2, 3, 1, 77 + seed), ref)


def test_errors_follow_the_reference():
    wm, tp, n_init = synth.make_hospital(2, 10, 3, seed=1)
    par = synth.tutorial_parameters(2)
    icp = np.full((n_init, 2), 0.2)
    with pytest.raises(IndexError):   # ward missing from total_patients_per_ward
        oracle.simulate_discrete_model(icp, icp, wm, tp[tp[:, 1] != 1], par, 2, 7, 1, 1)
    with pytest.raises(IndexError):   # too few parameter rows
        oracle.simulate_discrete_model(icp, icp, wm, tp, par[:1], 2, 7, 1, 1)
    with pytest.raises(RuntimeError):  # 7-column ward_matrix
        oracle.simulate_discrete_model(icp, icp, np.hstack([wm, wm[:, :1]]), tp, par, 2, 7, 1, 1)
    bad = wm.copy(); bad[0, 5] = wm.shape[0] + 5
    with pytest.raises(IndexError):   # next_day out of range
        oracle.simulate_discrete_model(icp, icp, bad, tp, par, 2, 7, 1, 1)


from tests import reference_cases  # noqa: E402


@pytest.mark.parametrize("case", reference_cases.ALL, ids=lambda f: f.__name__)
def test_reference_suite_on_the_oracle(case):
    case(oracle)
