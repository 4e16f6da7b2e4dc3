"""CPU-side tests of the product: the C ABI library loads and exports what include/amro_b200.h declares, the host
topology compiler, error classes, the summary arithmetic and the ensemble plumbing (gloo, world_size 2).
No GPU compute is called here; the product has no CPU fallback and must say so."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from abm_amro_b200 import capi, engine, ensemble, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_symbol_the_header_declares():
    hdr = open(os.path.join(ROOT, "include", "amro_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(amro_[a-z_]+)\s*\(", hdr)))
    assert declared == sorted(capi.EXPORTS)
    lib = capi.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.amro_version() == b"0.2.2"


def test_drop_in_module_has_the_reference_api():
    from abm_amro_b200 import amro
    for f in ["simulate_discrete_model", "progress_patients_probability_ward_1_timestep", "progress_patients_1_timestep",
              "total_positive_colonized", "total_positive_detected", "summary_of_total_positive", "simulate_and_collapse"]:
        assert callable(getattr(amro, f))
    assert amro.__version__ == "0.2.2"
    doc = amro.simulate_discrete_model.__doc__
    for kw in ["initial_colonized_probability", "initial_detected_probability", "ward_matrix", "total_patients_per_ward",
               "parameters", "time_to_detect: typing.SupportsInt | typing.SupportsIndex = 6" if "SupportsIndex" in doc else "time_to_detect",
               "testing_schedule_hospitalized", "testing_schedule_arrivals", "seed"]:
        assert kw.split(":")[0] in doc


def test_no_cpu_fallback():
    wm, tp, n_init = synth.make_hospital(2, 10, 3, seed=1)
    with engine.Topology(wm, tp, n_init, device=-1) as t:      # host-only handle
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            t.simulate(synth.tutorial_parameters(2), 0.2, 0.2, 2, 7, 1, 1)
    if capi.lib().amro_device_count() == 0:
        from abm_amro_b200 import amro
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            amro.simulate_and_collapse(wm, tp, 0.2, 0.2, n_init, .1, .2, .3, .4, .5, .6, 4, 2, 7, 1)


def _python_topology(wm, n_init):
    """Independent restatement of what the compiler must produce for a sorted hospital with one-day links."""
    R = wm.shape[0]
    day, wardc, nd = wm[:, 0].astype(int), wm[:, 1], wm[:, 5].astype(int)
    day_start = np.searchsorted(day, np.arange(day.max() + 2))
    seg_id = np.zeros(R, dtype=np.int64)
    seg_id[1:] = np.cumsum((day[1:] != day[:-1]) | (wardc[1:] != wardc[:-1]))
    day_seg_start = seg_id[np.minimum(day_start[:-1], R - 1)]
    meta_next = np.full(R, 0xFFFFFFFF, dtype=np.uint64)
    meta_nseg = np.full(R, 0xFFFFFFFF, dtype=np.uint64)
    has_pred = np.zeros(R, dtype=bool)
    src = np.nonzero(nd > 0)[0]
    meta_next[src] = nd[src] - day_start[day[nd[src]]]
    meta_nseg[src] = seg_id[nd[src]]
    has_pred[nd[src]] = True
    kind = np.where(has_pred, 1, np.where(np.arange(R) < n_init, 2, 0))
    info = (wm[:, 3].astype(np.uint64) << 31) | (kind.astype(np.uint64) << 29) | (seg_id - day_seg_start[day]).astype(np.uint64)
    return day_start, seg_id, info, meta_next, meta_nseg


def test_topology_compiler_on_a_synthetic_hospital():
    wm, tp, n_init = synth.make_hospital(5, 120, 25, seed=3)
    with engine.Topology(wm, tp, n_init, device=-1) as t:
        i = t.info
        assert (i.n_rows, i.n_rows_stepped, i.n_init, i.total_days, i.n_days_out) == (wm.shape[0], wm.shape[0], n_init, 24, 25)
        assert i.fast_eligible and i.identity_order and i.why_not_fast == b""
        a = t.debug_arrays()
    day_start, seg_id, info, nxt, nseg = _python_topology(wm, n_init)
    assert np.array_equal(a["order"], np.arange(wm.shape[0]))
    assert np.array_equal(a["day_start"], day_start)
    assert np.array_equal(a["seg_row_start"][:-1], np.nonzero(np.r_[True, seg_id[1:] != seg_id[:-1]])[0])
    assert np.array_equal(a["meta_info"].astype(np.uint64), info)
    assert np.array_equal(a["meta_next"].astype(np.uint64), nxt)
    assert np.array_equal(a["meta_nseg"].astype(np.uint64), nseg)
    # N of a ward-day is the FIRST matching row of total_patients_per_ward (abm_wards.cpp:242,249)
    tp2 = np.vstack([tp, tp * [1, 1, 7]])
    with engine.Topology(wm, tp2, n_init, device=-1) as t:
        assert np.array_equal(t.debug_arrays()["seg_N"], a["seg_N"])


@pytest.mark.parametrize("kind", ["sorted", "shuffled", "exotic"])
def test_threaded_compile_gives_the_arrays_of_the_serial_one(monkeypatch, kind):
    # hospitals of millions of rows are compiled by up to 16 threads (topology.cpp: blocks of rows / days); every array must be
    # what one thread produces -- sorted input, shuffled rows, and links with several sources per row (sequential fallback)
    wm, tp, n_init = synth.make_hospital(5, 300, 14, seed=21, p_split=0.1)
    rng = np.random.default_rng(4)
    R = wm.shape[0]
    if kind != "sorted":
        perm = rng.permutation(R)
        inv = np.empty(R, dtype=np.int64)
        inv[perm] = np.arange(R)
        wm = wm[perm].copy(order="F")
        nd = wm[:, 5].astype(np.int64)
        nd[nd >= 0] = inv[nd[nd >= 0]]
        wm[:, 5] = nd
    if kind == "exotic":
        idx = rng.integers(0, R, size=R // 25)
        wm[idx, 5] = rng.integers(0, R, size=idx.size)
        wm[rng.integers(0, R, size=5), 0] += 0.5
    got = {}
    for name, env in (("serial", {"AMRO_COMPILE_THREADS": "1"}), ("threads", {"AMRO_COMPILE_PAR_MIN_ROWS": "1", "AMRO_COMPILE_THREADS": "5"})):
        for k in ("AMRO_COMPILE_THREADS", "AMRO_COMPILE_PAR_MIN_ROWS"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        with engine.Topology(wm, tp, n_init, device=-1) as t:
            i = t.info
            got[name] = (t.debug_arrays(), (i.n_rows_stepped, i.n_segments, i.max_rows_per_day, i.fast_eligible, i.identity_order, bytes(i.why_not_fast)))
    assert got["serial"][1] == got["threads"][1]
    for k, v in got["serial"][0].items():
        w = got["threads"][0][k]
        assert (v is None) == (w is None) and (v is None or np.array_equal(v, w)), k


def test_topology_compiler_sorts_like_the_reference_visits():
    wm, tp, n_init = synth.make_hospital(3, 30, 6, seed=5)
    g = np.random.default_rng(0)
    perm = np.r_[np.arange(n_init), n_init + g.permutation(wm.shape[0] - n_init)]   # initial rows must stay first (:365)
    inv = np.empty_like(perm); inv[perm] = np.arange(perm.size)
    sh = wm[perm].copy()
    m = sh[:, 5] > 0
    sh[m, 5] = inv[sh[m, 5].astype(int)]
    sh = np.asfortranarray(sh)
    with engine.Topology(sh, tp, n_init, device=-1) as t:
        a = t.debug_arrays()
        assert not t.info.identity_order
    order = a["order"]
    key = np.stack([sh[order, 0], sh[order, 1]], 1)
    assert np.all((key[1:, 0] > key[:-1, 0]) | ((key[1:, 0] == key[:-1, 0]) & (key[1:, 1] >= key[:-1, 1])))
    same = (key[1:] == key[:-1]).all(1)
    assert np.all(order[1:][same] > order[:-1][same])          # stable: original row order inside a ward-day


def test_push_lists_keep_the_last_writer_of_every_target():
    # abm_wards.cpp:402-407: after day d, its rows with next_day > 0 write into row trunc(next_day) in ascending row order,
    # so the LAST source of a target wins; the compiled per-day push lists hold exactly those winners, ascending by source
    wm, tp, n_init = synth.make_hospital(4, 120, 10, seed=3)
    g = np.random.default_rng(0)
    w2 = wm.copy(); R = w2.shape[0]
    idx = g.choice(R, 300, replace=False)
    w2[idx, 5] = g.integers(1, R, 300) + g.random(300) * 0.9      # duplicate, backward and fractional (truncated) targets
    with engine.Topology(np.asfortranarray(w2), tp, n_init, device=-1) as t:
        a = t.debug_arrays()
        assert not t.info.fast_eligible
    exp_src, exp_dst = [], []
    for d in range(int(tp[:, 0].max()) + 1):
        winner = {}
        for r in np.nonzero((w2[:, 0] == d) & (w2[:, 5] > 0))[0]:
            winner[int(w2[r, 5])] = int(r)
        pairs = sorted((r, tgt) for tgt, r in winner.items())
        exp_src += [r for r, _ in pairs]; exp_dst += [tgt for _, tgt in pairs]
    assert np.array_equal(a["push_src"], exp_src) and np.array_equal(a["push_dst"], exp_dst)


def test_fast_path_eligibility_reasons():
    wm, tp, n_init = synth.make_hospital(3, 30, 6, seed=5, p_split=0.3)
    with engine.Topology(wm, tp, n_init, device=-1) as t:
        assert t.info.fast_eligible                                             # weights 1 and 1/2: pressure counted in halves
    w2 = wm.copy(); w2[w2[:, 4] == 0.5, 4] = 0.3
    with engine.Topology(w2, tp, n_init, device=-1) as t:
        assert not t.info.fast_eligible and b"weights" in t.info.why_not_fast
    wm, tp, n_init = synth.make_hospital(3, 30, 6, seed=5)
    w2 = wm.copy(); w2[0, 3] = 0.5
    with engine.Topology(w2, tp, n_init, device=-1) as t:
        assert b"is_new" in t.info.why_not_fast
    w2 = wm.copy(); src = np.nonzero(w2[:, 5] > 0)[0][-1]; w2[src, 5] = 3       # link back to day 0
    with engine.Topology(w2, tp, n_init, device=-1) as t:
        assert b"earlier day" in t.info.why_not_fast
    w2 = wm.copy()
    d0 = np.nonzero((w2[:, 0] == 0) & (w2[:, 5] > 0))[0][0]
    tgt2 = np.nonzero((w2[:, 0] == 2) & (w2[:, 3] == 1))[0][-1]                  # an arrival: nobody else writes it
    w2[d0, 5] = tgt2                                                              # day 0 -> day 2
    with engine.Topology(w2, tp, n_init, device=-1) as t:
        assert b"skip a day" in t.info.why_not_fast
    with engine.Topology(wm, tp[tp[:, 0] < 4], n_init, device=-1) as t:          # days 4-5 never stepped
        assert b"never stepped" in t.info.why_not_fast and t.info.n_rows_stepped < t.info.n_rows


def test_error_classes_follow_the_reference():
    wm, tp, n_init = synth.make_hospital(2, 10, 3, seed=1)
    with pytest.raises(IndexError):
        engine.Topology(wm, tp[tp[:, 1] != 1], n_init, device=-1)
    with pytest.raises(RuntimeError):
        engine.Topology(np.hstack([wm, wm[:, :1]]), tp, n_init, device=-1)
    bad = wm.copy(); bad[0, 5] = wm.shape[0] + 5
    with pytest.raises(IndexError):
        engine.Topology(bad, tp, n_init, device=-1)
    with pytest.raises(IndexError):
        engine.Topology(wm, tp, wm.shape[0] + 1, device=-1)
    with pytest.raises(RuntimeError):
        engine.Topology(wm, tp[:0], n_init, device=-1)


def test_summary_of_total_positive_is_host_arithmetic(golden):
    import ctypes
    z = golden("summary")
    pos = np.asfortranarray(z["col"]); q = np.ascontiguousarray(z["q"])
    out = np.empty((pos.shape[0], 3 + q.size), order="F")
    capi.check(capi.lib().amro_summary_of_total_positive(pos.ctypes.data, pos.shape[0], pos.shape[1], pos.shape[0],
                                                         q.ctypes.data, q.size, out.ctypes.data))
    assert np.array_equal(out, z["summ"])
    from abm_amro_b200 import amro
    assert np.array_equal(amro.summary_of_total_positive(z["col"], z["q"]), z["summ"])


def test_synthetic_hospital_invariants():
    wm, tp, n_init = synth.make_hospital(4, 50, 10, seed=9)
    assert n_init == 50 and wm.shape == (500, 6) and np.all(wm[:n_init, 3] == 1)
    nd = wm[:, 5].astype(int)
    src = np.nonzero(nd > 0)[0]
    assert np.all(wm[nd[src], 0] == wm[src, 0] + 1) and np.all(wm[nd[src], 2] == wm[src, 2])   # same patient, next day
    assert np.all(wm[nd[src], 3] == 0)
    tot = {(d, w): n for d, w, n in tp}
    for d in range(10):
        for w in np.unique(wm[wm[:, 0] == d, 1]):
            assert tot[(d, w)] == np.sum((wm[:, 0] == d) & (wm[:, 1] == w))


def test_shard_sims_tiles_the_ensemble():
    for total in [1, 7, 8, 1000, 1001, 65536]:
        for world in [1, 2, 3, 8]:
            blocks = [ensemble.shard_sims(total, world, r) for r in range(world)]
            assert all(o % 32 == 0 for o, _ in blocks)   # a group of 32 simulations shares the plane words of the daily draws
            assert sum(n for _, n in blocks) == total
            live = [b for b in blocks if b[1] > 0]
            assert live[0][0] == 0 and all(live[i][0] + live[i][1] == live[i + 1][0] for i in range(len(live) - 1))


_GLOO_WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from abm_amro_b200 import ensemble, synth
from oracle import oracle
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
S = 80
wm, tp, n_init = synth.make_hospital(3, 40, 10, seed=2)
par = synth.tutorial_parameters(S)
icp = np.full((n_init, S), 0.2); idp = np.full((n_init, S), 0.3)
def counts(off, n):   # the per-rank compute stand-in: the oracle's Philox stream, offset by the GLOBAL simulation index
    m = oracle.simulate_discrete_model(icp[:, off:off+n], idp[:, off:off+n], wm, tp, par[off:off+n], 2, 7, 1, 5,
                                       rng=oracle.Rng(oracle.RNG_PHILOX, sim_offset=off))
    return oracle.total_positive_colonized(m, n), oracle.total_positive_detected(m, n)
off, n = ensemble.shard_sims(S, world, rank)
c, d = counts(off, n)
mom = ensemble.day_moments(torch.from_numpy(np.ascontiguousarray(c.T)), torch.from_numpy(np.ascontiguousarray(d.T)))
got = ensemble.collapse(mom, S).numpy()
# the fused run hands over the transpose of a [4 x n_days] buffer (bench.py): a non-contiguous view must work too
got_t = ensemble.collapse(mom.t().contiguous().t(), S).numpy()
assert np.array_equal(got, got_t)
cf, df = counts(0, S)
want = np.stack([cf.mean(1), df.mean(1), cf.std(1), df.std(1)], 1)
assert np.allclose(got, want, rtol=0, atol=1e-9), np.abs(got - want).max()
if rank == 0: print("gloo ok")
dist.destroy_process_group()
"""


def test_ensemble_collapse_world_size_2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "gloo ok" in outs[0][0]


def test_synthetic_hospital_tables_are_a_valid_fast_hospital():
    # amro_hospital_tables: the "bed model" hospital the device generator builds, as the reference's float64 tables
    spec = engine.hospital_spec(6, 120, 9, seed=11, p_discharge=0.2, p_transfer=0.3)
    wm, tp, n_init = engine.hospital_tables(spec)
    wm2, tp2, _ = engine.hospital_tables(spec)
    assert np.array_equal(wm, wm2) and np.array_equal(tp, tp2) and n_init == 120 and wm.shape == (120 * 9, 6)
    assert not np.array_equal(wm, engine.hospital_tables(engine.hospital_spec(6, 120, 9, seed=12, p_discharge=0.2, p_transfer=0.3))[0])
    nd = wm[:, 5].astype(np.int64)
    src = np.nonzero(nd > 0)[0]
    assert np.all(wm[nd[src], 0] == wm[src, 0] + 1) and np.all(wm[nd[src], 2] == wm[src, 2])     # the same patient, the next day
    assert len(np.unique(nd[src])) == len(src)                                                      # one predecessor at most
    targets = np.zeros(wm.shape[0], dtype=bool); targets[nd[src]] = True
    assert np.array_equal(wm[:, 3] == 1, ~targets)                                                  # is_new <=> no predecessor
    assert np.all(np.diff(wm[:, 0] * 1000 + wm[:, 1]) >= 0) and np.all(wm[:, 4] == 1.0)            # rows in (day, ward) order
    for d, w, c in tp:
        assert np.sum((wm[:, 0] == d) & (wm[:, 1] == w)) == c
    stay = len(src) / (120 * 8)
    assert abs(stay - 0.8) < 0.05                                                                   # 1 - p_discharge of the records have a successor
    with engine.Topology(wm, tp, n_init, device=-1) as topo:
        assert topo.info.fast_eligible and topo.info.identity_order
    with pytest.raises(ValueError):
        engine.hospital_tables(engine.hospital_spec(0, 10, 3))


_SWEEP_WORKER = r"""
import os, sys, types, numpy as np, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from abm_amro_b200 import capi, inference
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
class FakeTopology:   # stands in for the GPU: the distance of a particle is a known function of its parameters and global index
    def simulate(self, par, icp, idp, ttd, tsh, tsa, seed, sim_offset=0, want_counts=True, observed_colonized=None, observed_detected=None):
        n = par.shape[0]
        d = ((par[:, 1] - 0.6) ** 2) * 100 + (np.arange(sim_offset, sim_offset + n) % 7) * 1e-6
        return types.SimpleNamespace(distance=d, path_used=capi.PATH_FAST, kernel_ms=1.0)
n = 1000
res = inference.abc_sweep(FakeTopology(), np.zeros(3), None, n_particles=n, keep=0.05)
par = inference.UniformPrior().sample(n, 100)
want = ((par[:, 1] - 0.6) ** 2) * 100 + (np.arange(n) % 7) * 1e-6
assert np.array_equal(res.distance, want) and np.array_equal(res.particles, par)
assert np.array_equal(res.accepted, np.argsort(want, kind="stable")[:50])
assert abs(res.posterior_mean[1] - 0.6) < 0.02
if rank == 0: print("sweep ok")
dist.destroy_process_group()
"""


def test_abc_sweep_shards_particles_world_size_2_gloo(tmp_path):
    script = tmp_path / "s.py"
    script.write_text(_SWEEP_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29537", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "sweep ok" in outs[0][0]
