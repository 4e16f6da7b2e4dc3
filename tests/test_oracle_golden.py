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
