"""TEST INFRASTRUCTURE -- ctypes front-end of the C oracle (oracle/amro_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; nothing under abm_amro_b200/ does.  Function names and argument order mirror the
reference's Python API (/root/reference/src/amro/amro.cpp:18-318) so parity tests read like the
reference's own tests.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

RNG_GLIBC, RNG_STREAM, RNG_TENSOR, RNG_PHILOX = 0, 1, 2, 3

_dp = C.POINTER(C.c_double)
_u64 = C.c_uint64


class _Rng(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32),
                ("stream", _dp), ("stream_len", _u64), ("stream_pos", _u64),
                ("u_col", _dp), ("u_det", _dp), ("u_ld", _u64),
                ("u_icol", _dp), ("u_idet", _dp), ("ui_ld", _u64),
                ("seed", _u64), ("sim_offset", _u64), ("ward_day", _u64), ("n_draws", _u64),
                ("rec_col", _dp), ("rec_det", _dp), ("rec_ld", _u64),
                ("rec_icol", _dp), ("rec_idet", _dp), ("reci_ld", _u64)]


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (oracle/Makefile `oracle` target)."""
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("amro_oracle.c", "amro_oracle.h")):
        subprocess.run(["make", "-C", _HERE, "oracle"], check=True, stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.amro_oracle_last_error.restype = C.c_char_p
        _lib.amro_oracle_threshold.restype = C.c_uint64
        _lib.amro_oracle_threshold.argtypes = [C.c_double]
        _lib.amro_oracle_philox_u32.restype = C.c_uint32
        _lib.amro_oracle_philox_u32.argtypes = [_u64, _u64, _u64, C.c_uint32]
        _lib.amro_oracle_philox_chunk.restype = C.c_uint32
        _lib.amro_oracle_philox_chunk.argtypes = [_u64, _u64, _u64, C.c_uint32]
        _lib.amro_oracle_threshold30.restype = C.c_uint32
        _lib.amro_oracle_threshold30.argtypes = [C.c_double]
        _lib.amro_oracle_bernoulli14.restype = C.c_uint32
        _lib.amro_oracle_bernoulli14.argtypes = [C.c_double, C.c_uint32]
        _lib.amro_oracle_total_positive.restype = C.c_int64
        _lib.amro_oracle_simulate_and_collapse.restype = C.c_int64
    return _lib


def _raise(rc):
    msg = lib().amro_oracle_last_error().decode()
    raise (IndexError if rc == -1 else RuntimeError)(msg)


def _f(a):
    """float64, 2-D, Fortran order (1-D -> column), like the reference's carma conversion."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    if a.ndim != 2:
        raise RuntimeError("CARMA: Number of dimensions must be 1 <= ndim <= 2")
    return np.asfortranarray(a)


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _mat(a):
    return _p(a), _u64(a.shape[0]), _u64(a.shape[1]), _u64(a.shape[0])


class Rng:
    """Uniform source handed to the oracle."""

    def __init__(self, kind=RNG_GLIBC, stream=None, u_col=None, u_det=None, u_icol=None, u_idet=None,
                 seed=0, sim_offset=0, record_shape=None, record_init_shape=None):
        self.s = _Rng()
        self.s.kind = kind
        self._keep = []
        if stream is not None:
            stream = np.ascontiguousarray(stream, dtype=np.float64)
            self._keep.append(stream)
            self.s.stream, self.s.stream_len = _p(stream), stream.size
        for name, arr in (("u_col", u_col), ("u_det", u_det)):
            if arr is not None:
                arr = _f(arr); self._keep.append(arr)
                setattr(self.s, name, _p(arr)); self.s.u_ld = arr.shape[0]
        for name, arr in (("u_icol", u_icol), ("u_idet", u_idet)):
            if arr is not None:
                arr = _f(arr); self._keep.append(arr)
                setattr(self.s, name, _p(arr)); self.s.ui_ld = arr.shape[0]
        self.s.seed = seed & 0xFFFFFFFFFFFFFFFF
        self.s.sim_offset = sim_offset
        self.rec_col = self.rec_det = self.rec_icol = self.rec_idet = None
        if record_shape is not None:
            self.rec_col = np.full(record_shape, np.nan, order="F")
            self.rec_det = np.full(record_shape, np.nan, order="F")
            self.s.rec_col, self.s.rec_det, self.s.rec_ld = _p(self.rec_col), _p(self.rec_det), record_shape[0]
        if record_init_shape is not None:
            self.rec_icol = np.full(record_init_shape, np.nan, order="F")
            self.rec_idet = np.full(record_init_shape, np.nan, order="F")
            self.s.rec_icol, self.s.rec_idet, self.s.reci_ld = _p(self.rec_icol), _p(self.rec_idet), record_init_shape[0]

    @property
    def n_draws(self):
        return int(self.s.n_draws)

    @property
    def stream_pos(self):
        return int(self.s.stream_pos)


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
    lib().amro_oracle_philox4x32_10(c, k, o)
    return [int(x) for x in o]


def threshold(p: float) -> int:
    return int(lib().amro_oracle_threshold(float(p)))


def philox_u32(seed, row, sim, kind) -> int:
    return int(lib().amro_oracle_philox_u32(seed & 0xFFFFFFFFFFFFFFFF, row, sim, kind))


def philox_chunk(seed, index, sim, stream) -> int:
    """16-bit chunk of (index, global simulation, stream): the daily draw (stream 0, & 0x3FFF) and the ward-day dither (stream 6)."""
    return int(lib().amro_oracle_philox_chunk(seed & 0xFFFFFFFFFFFFFFFF, index, sim, stream))


def threshold30(p: float) -> int:
    return int(lib().amro_oracle_threshold30(float(p)))


def bernoulli14(p: float, r16: int) -> int:
    """Dithered 14-bit threshold of the daily step: (T30(p) + r16) >> 16."""
    return int(lib().amro_oracle_bernoulli14(float(p), int(r16)))


def progress_patients_probability_ward_1_timestep(ward_matrix, total_patients, parameters, n_sims,
                                                  time_to_detect, detect_today_hospitalized, detect_today_arrival,
                                                  rng: Rng | None = None, row_ids=None, return_p=False):
    wm = _f(ward_matrix).copy(order="F"); par = _f(parameters)
    rng = rng or Rng()
    ids = None if row_ids is None else np.ascontiguousarray(row_ids, dtype=np.uint64)
    n_ids = wm.shape[0] if ids is None else int(ids.max()) + 1 if ids.size else 0
    p_out = np.full((n_ids, max(int(n_sims), 1)), np.nan, order="F") if return_p else None
    rc = lib().amro_oracle_ward_step(*_mat(wm), C.c_double(total_patients), *_mat(par), _u64(n_sims),
                                     C.c_int(time_to_detect), C.c_int(bool(detect_today_hospitalized)),
                                     C.c_int(bool(detect_today_arrival)), C.byref(rng.s),
                                     None if ids is None else ids.ctypes.data_as(C.POINTER(_u64)),
                                     _p(p_out), _u64(n_ids))
    if rc:
        _raise(rc)
    return (wm, p_out) if return_p else wm


def progress_patients_1_timestep(ward_matrix, total_patients_per_ward, parameters, n_sims, time_to_detect,
                                 detect_today_hospitalized, detect_today_arrival, rng: Rng | None = None,
                                 return_p=False):
    wm = _f(ward_matrix).copy(order="F"); tp = _f(total_patients_per_ward); par = _f(parameters)
    rng = rng or Rng()
    p_out = np.full((wm.shape[0], max(int(n_sims), 1)), np.nan, order="F") if return_p else None
    rc = lib().amro_oracle_day_step(*_mat(wm), *_mat(tp), *_mat(par), _u64(n_sims), C.c_int(time_to_detect),
                                    C.c_int(bool(detect_today_hospitalized)), C.c_int(bool(detect_today_arrival)),
                                    C.byref(rng.s), None, _p(p_out), _u64(wm.shape[0]))
    if rc:
        _raise(rc)
    return (wm, p_out) if return_p else wm


def simulate_discrete_model(initial_colonized_probability, initial_detected_probability, ward_matrix,
                            total_patients_per_ward, parameters, time_to_detect=6,
                            testing_schedule_hospitalized=7, testing_schedule_arrivals=8, seed=9,
                            rng: Rng | None = None, return_p=False):
    icp = _f(initial_colonized_probability); idp = _f(initial_detected_probability)
    wm = _f(ward_matrix); tp = _f(total_patients_per_ward); par = _f(parameters)
    rng = rng or Rng()
    R, S = wm.shape[0], icp.shape[1]
    out = np.empty((R, wm.shape[1] + 2 * S), order="F")
    p_out = np.full((R, max(S, 1)), np.nan, order="F") if return_p else None
    rc = lib().amro_oracle_simulate(*_mat(icp), *_mat(idp), *_mat(wm), *_mat(tp), *_mat(par),
                                    C.c_int(time_to_detect), C.c_int(testing_schedule_hospitalized),
                                    C.c_int(testing_schedule_arrivals), C.c_int(seed), C.byref(rng.s),
                                    _p(out), _p(p_out), _u64(R))
    if rc:
        _raise(rc)
    return (out, p_out) if return_p else out


def _total(m, col_init, col_end, mode):
    m = _f(m)
    days = lib().amro_oracle_total_positive(*_mat(m), _u64(col_init), _u64(col_end), C.c_int(mode), None)
    if days < 0:
        _raise(days)
    out = np.empty((days, col_end - col_init + 1), order="F")
    rc = lib().amro_oracle_total_positive(*_mat(m), _u64(col_init), _u64(col_end), C.c_int(mode), _p(out))
    if rc < 0:
        _raise(rc)
    return out


def total_positive_colonized(model_colonized, n_sims=2):
    return _total(model_colonized, 6, 6 + n_sims - 1, 1)


def total_positive_detected(model_colonized, n_sims=2):
    m = _f(model_colonized)
    return _total(m, 6 + n_sims, m.shape[1] - 1, 0)


def summary_of_total_positive(model_positive, quantiles):
    pos = _f(model_positive); q = np.ascontiguousarray(np.asarray(quantiles, dtype=np.float64).ravel())
    out = np.empty((pos.shape[0], 3 + q.size), order="F")
    rc = lib().amro_oracle_summary(_p(pos), _u64(pos.shape[0]), _u64(pos.shape[1]), _u64(pos.shape[0]),
                                   _p(q), _u64(q.size), _p(out))
    if rc:
        _raise(rc)
    return out


def simulate_and_collapse(ward_matrix, total_patients_per_ward, initial_colonized_probability,
                          initial_detected_probability, initial_patients, alpha, beta, gamma, rho_hospital,
                          alpha2, rho_imported, n_sims=100, time_to_detect=None,
                          testing_schedule_hospitalized=None, testing_schedule_arrivals=None, seed=2384235,
                          rng: Rng | None = None):
    wm = _f(ward_matrix); tp = _f(total_patients_per_ward)
    rng = rng or Rng()
    args = [*_mat(wm), *_mat(tp), C.c_double(initial_colonized_probability), C.c_double(initial_detected_probability),
            _u64(initial_patients), C.c_double(alpha), C.c_double(beta), C.c_double(gamma), C.c_double(rho_hospital),
            C.c_double(alpha2), C.c_double(rho_imported), _u64(n_sims), _u64(time_to_detect),
            _u64(testing_schedule_hospitalized), _u64(testing_schedule_arrivals), C.c_int(seed), C.byref(rng.s)]
    days = lib().amro_oracle_simulate_and_collapse(*args, None)
    if days < 0:
        _raise(days)
    out = np.empty((days, 5), order="F")
    rc = lib().amro_oracle_simulate_and_collapse(*args, _p(out))
    if rc < 0:
        _raise(rc)
    return out
