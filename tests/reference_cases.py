"""The reference's own known-answer cases (/root/reference/tests/test_model.py:4-362), restated data-driven so the
same checks run against the CPU oracle and against the CUDA module `abm_amro_b200.amro`.  Each case picks degenerate
probabilities (0 or 1) so the stochastic step is deterministic.  `m` is anything with the reference's API."""
import numpy as np

INF = np.inf


def ward(day, ward_id, mrn, new, nxt, *state_cols):
    n = len(mrn)
    cols = [np.broadcast_to(day, n), np.broadcast_to(ward_id, n), mrn, new, np.ones(n), nxt] + list(state_cols)
    return np.asfortranarray(np.array(cols, dtype=np.float64).T)


def params(alpha, beta, gamma, rho_h, alpha2, rho_i):
    return np.array([[alpha, beta, gamma, rho_h, alpha2, rho_i]], dtype=np.float64)


MRN10 = [52, 53, 17, 20, 87, 99, 12, 43, 59, 98]
MRN6 = [52, 53, 17, 200, 87, 99]
D10 = [INF, 4, INF, INF, 1, INF, INF, 0, 0, INF]


def case_clearance_100(m):  # test_model.py:4-30
    w = ward(0, 1, MRN10, [0] * 6 + [1] * 4, [-1] * 10, [0, 1, 0, 0, 1, 1, 0, 0, 1, 1], D10)
    out = m.progress_patients_probability_ward_1_timestep(w, 100, params(1, 0, 0, .12, 1, .3), 1, 0, 1, 1)
    assert np.all(out[:, 6] == 0) and np.all(out[:, 7] == INF)


def case_clearance_0(m):  # :32-58
    w = ward(0, 1, MRN10, [0] * 6 + [1] * 4, [-1] * 10, [1] * 10, D10)
    out = m.progress_patients_probability_ward_1_timestep(w, 100, params(0, 0, 0, .12, 0, .1), 1, 0, 1, 1)
    assert np.all(out[:, 6] == 1)


def case_rho_hospital_100(m):  # :60-86
    w = ward(0, 1, MRN6, [0, 1, 1, 0, 0, 0], [-1] * 6, [0, 0, 0, 1, 1, 1], [INF] * 6)
    out = m.progress_patients_probability_ward_1_timestep(w, 6, params(0, 0, 0, 1, 0, 0), 1, 0, 1, 1)
    assert np.all(out[3:, 7] == 0)


def case_rho_hospital_0(m):  # :88-114
    w = ward(0, 1, MRN6, [0, 1, 1, 0, 0, 0], [-1] * 6, [0, 1, 1, 0, 1, 1], [INF] * 6)
    out = m.progress_patients_probability_ward_1_timestep(w, 6, params(0, 0, 0, 0, 0, 0), 1, 0, 1, 1)
    assert np.all(out[:, 7] == INF)


def case_rho_imported_100(m):  # :116-142
    w = ward(0, 1, MRN6, [1, 1, 0, 0, 0, 0], [-1] * 6, [0] * 6, [INF] * 6)
    out = m.progress_patients_probability_ward_1_timestep(w, 6, params(0, 0, 1, 0, 0, 1), 1, 0, 1, 1)
    assert np.all((out[:, 6] == 1) == (out[:, 7] <= 0)) and np.all(out[:, 6] == out[:, 3])


def case_rho_imported_0(m):  # :144-170
    w = ward(0, 1, MRN6, [1, 1, 1, 1, 0, 0], [-1] * 6, [0] * 6, [INF] * 6)
    out = m.progress_patients_probability_ward_1_timestep(w, 6, params(0, 0, 1, 0, 0, 0), 1, 0, 1, 1)
    assert np.all(out[:, 7] == INF) and np.all(out[:, 6] == out[:, 3])


def case_gamma_1_arrivals_infected(m):  # :172-200 -- relies on the literal defaults tsh=7, tsa=8, seed=9
    tp = np.array([[0, 1, 20]], dtype=np.float64)
    z = np.zeros((6, 1))
    w = ward(0, 1, MRN6, [1, 1, 0, 0, 1, 0], [-1] * 6)
    run = m.simulate_discrete_model(z, z, w, tp, params(.1, .2, 1, .3, .2, .3), 10)
    assert np.all(run[:, 3] == run[:, 6])


def case_detected_persistence(m):  # :202-231
    tp = np.array([[0, 1, 20], [1, 1, 10], [2, 1, 12]], dtype=np.float64)
    init = np.array([[1, 0, 1, 0, 0, 1]], dtype=np.float64).T
    mrn = [52, 53, 17, 200, 87, 99, 52, 53, 17, 200, 99, 52, 53, 17, 200, 99]
    w = ward([0] * 6 + [1] * 5 + [2] * 5, 1, mrn, [1, 1, 0, 0, 1, 0] + [0] * 10,
             [6, 7, 8, 9, -1, 10, 11, 12, 13, 14, 15, -1, -1, -1, -1, -1])
    run = m.simulate_discrete_model(init, init, w, tp, params(0, 0, 0, 1, 0, 1), 0, 1, 1)
    assert np.all((run[:, 6] == 1) == (run[:, 7] <= 0))


def case_initial_detected(m):  # :233-261
    tp = np.array([[0, 1, 20]], dtype=np.float64)
    init_c = np.array([[1, 0, 1, 0, 0, 1]], dtype=np.float64).T
    w = ward(0, 1, MRN6, [1, 1, 0, 0, 1, 0], [-1] * 6)
    run = m.simulate_discrete_model(init_c, np.ones((6, 1)), w, tp, params(.1, .2, .3, 1, .2, 1), 0, 1, 1)
    assert np.all((run[:, 7] <= 0) == (run[:, 6] == 1))


def _four_model_matrix(cols):
    return ward(0, 1, MRN6, [1, 1, 0, 0, 1, 0], [-1] * 6, *cols)


def case_sum_colonized(m):  # :263-277 -- relies on the default n_sims = 2
    w = _four_model_matrix([[0, 0, 0, 1, 0, 0], [1, 0, 1, 1, 0, 0], [1, 1, 1, 1, 0, 0], [1, 0, 1, 1, 0, 0]])
    assert np.all(m.total_positive_colonized(w) == np.array([1, 3]))


def case_sum_detected(m):  # :279-293
    w = _four_model_matrix([[INF, INF, INF, -1, INF, INF], [-3, INF, -2, -1, INF, INF], [-1, 0, 0, -1, INF, INF],
                            [-1, INF, -1, -4, INF, INF]])
    assert np.all(m.total_positive_detected(w) == np.array([4, 3]))


def _two_day_matrix():
    mrn = [52, 53, 17, 200, 87, 99, 52, 53, 17, 20]
    return ward([0] * 6 + [1] * 4, 1, mrn, [1, 1, 0, 0, 1, 0, 0, 1, 1, 0], [6, 7, 8, -1, -1, -1, -1, -1, -1, 1],
                [0, 0, 0, 1, 0, 0, 1, 0, 1, 0], [1, 0, 1, 1, 0, 0, 1, 0, 1, 0], [1, 1, 1, 1, 0, 0, 1, 0, 1, 0],
                [1, 0, 0, 1, 0, 0, 1, 0, 1, 0])


def case_summary_mean_sd(m):  # :295-331
    pos = m.total_positive_colonized(_two_day_matrix(), 4)
    res = m.summary_of_total_positive(pos, np.array([0.2, 0.9]))
    assert np.all(np.mean(pos, axis=1) == res[:, 1]) and np.all(np.std(pos, axis=1) == res[:, 2])


def case_seed_equal(m):  # :333-362
    tp = np.array([[0, 1, 20], [1, 1, 12]], dtype=np.float64)
    ic = np.array([[0, 1, 0, 0, 0, 0]], dtype=np.float64).T
    idt = np.array([[0, .5, 0, 0, 0, 0]], dtype=np.float64).T
    w = _two_day_matrix()[:, :6]
    a = m.simulate_discrete_model(ic, idt, w, tp, params(.1, .2, .5, .8, .2, .3), 1, 1, 1, 34279)
    b = m.simulate_discrete_model(ic, idt, w, tp, params(.1, .2, .5, .8, .2, .3), 1, 1, 1, 34279)
    assert np.all(a == b)


ALL = [v for k, v in sorted(globals().items()) if k.startswith("case_")]
