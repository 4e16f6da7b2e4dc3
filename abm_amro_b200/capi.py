"""ctypes binding of the C ABI (include/amro_b200.h) -- exactly the symbols the header declares.

The shared library is built in-tree by ``abm_amro_b200/csrc/Makefile`` (``__graft_entry__.build()``).
Loading fails loudly when it is missing: there is no CPU fallback on the product path.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AMRO_B200_LIB") or os.path.join(_PKG, "lib", "libamro_b200.so")  # env override: tuning variants

AMRO_OK, AMRO_EINDEX, AMRO_ERUNTIME, AMRO_EINVAL, AMRO_ECUDA = 0, 1, 2, 3, 4
RNG_PHILOX, RNG_REPLAY = 0, 1
PATH_AUTO, PATH_FAST, PATH_GENERAL = 0, 1, 2
KERNEL_NONE, KERNEL_CODES, KERNEL_BITS, KERNEL_FLOAT64 = 0, 1, 2, 3
BARRIER_CTA, BARRIER_GLOBAL, BARRIER_CLUSTER = 0, 1, 2

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_i64 = C.c_int64


class TopologyInfo(C.Structure):
    _fields_ = [("n_rows", _i64), ("n_rows_stepped", _i64), ("n_init", _i64), ("total_days", _i64),
                ("n_days_out", _i64), ("n_segments", _i64), ("max_rows_per_day", _i64),
                ("fast_eligible", C.c_int32), ("ring_eligible", C.c_int32), ("identity_order", C.c_int32),
                ("device", C.c_int32), ("compile_ms", C.c_double), ("why_not_fast", C.c_char * 128)]


class HospitalSpec(C.Structure):
    _fields_ = [("n_wards", _i64), ("census", _i64), ("n_days", _i64), ("seed", _i64),
                ("p_discharge", C.c_double), ("p_transfer", C.c_double)]


class SimArgs(C.Structure):
    _fields_ = [("n_sims", _i64), ("sim_offset", _i64), ("seed", _i64),
                ("time_to_detect", C.c_int32), ("testing_schedule_hospitalized", C.c_int32),
                ("testing_schedule_arrivals", C.c_int32),
                ("rng_mode", C.c_int32), ("path", C.c_int32), ("keep_trajectory", C.c_int32),
                ("pointers_on_device", C.c_int32), ("async_", C.c_int32),
                ("parameters", C.c_void_p), ("p_rows", _i64), ("p_ld", _i64),
                ("icp", C.c_void_p), ("icp_rs", _i64), ("icp_cs", _i64),
                ("idp", C.c_void_p), ("idp_rs", _i64), ("idp_cs", _i64),
                ("u_col", C.c_void_p), ("u_det", C.c_void_p), ("u_ld", _i64),
                ("u_icol", C.c_void_p), ("u_idet", C.c_void_p), ("ui_ld", _i64),
                ("counts_colonized", C.c_void_p), ("counts_detected", C.c_void_p),
                ("state_out", C.c_void_p), ("so_ld", _i64),
                ("p_out", C.c_void_p), ("day_moments", C.c_void_p), ("pressure_out", C.c_void_p),
                ("observed_colonized", C.c_void_p), ("observed_detected", C.c_void_p), ("distance", C.c_void_p),
                ("quantiles", C.c_void_p), ("n_quantiles", _i64), ("summary_colonized", C.c_void_p), ("summary_detected", C.c_void_p),
                ("stream", C.c_void_p),
                ("path_used", C.c_int32), ("launches", C.c_int32), ("kernel_ms", C.c_float), ("total_ms", C.c_float),
                ("kernel_kind", C.c_int32), ("grid", C.c_int32), ("ctas_per_team", C.c_int32), ("team_barrier", C.c_int32)]


# every entry point include/amro_b200.h declares (tests/test_host.py checks this list against the header)
EXPORTS = ["amro_last_error", "amro_version", "amro_device_count", "amro_topology_create", "amro_topology_destroy",
           "amro_topology_get_info", "amro_topology_generate", "amro_hospital_tables", "amro_topology_kernel_times", "amro_topology_cache_clear", "amro_topology_cache_stats", "amro_simulate", "amro_simulate_discrete_model", "amro_simulate_and_collapse", "amro_simulate_and_collapse_alloc",
           "amro_progress_patients", "amro_total_positive", "amro_summary_of_total_positive", "amro_summary_of_total_positive_device",
           "amro_day_moments_peers"]

_lib = None


def lib():
    """The loaded shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(amro_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.amro_last_error.restype = C.c_char_p
        L.amro_version.restype = C.c_char_p
        L.amro_topology_create.argtypes = [C.c_void_p, _i64, _i64, _i64, C.c_void_p, _i64, _i64, _i64, _i64, C.c_int,
                                           C.POINTER(C.c_void_p)]
        L.amro_topology_destroy.argtypes = [C.c_void_p]
        L.amro_topology_destroy.restype = None
        L.amro_topology_get_info.argtypes = [C.c_void_p, C.POINTER(TopologyInfo)]
        L.amro_simulate.argtypes = [C.c_void_p, C.POINTER(SimArgs)]
        L.amro_topology_kernel_times.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
        L.amro_topology_generate.argtypes = [C.POINTER(HospitalSpec), C.c_int, C.POINTER(C.c_void_p)]
        L.amro_hospital_tables.argtypes = [C.POINTER(HospitalSpec), C.POINTER(_i64), C.POINTER(_i64), C.c_void_p, C.c_void_p]
        L.amro_topology_debug_device_arrays.argtypes = [C.c_void_p] * 8
        L.amro_topology_cache_stats.argtypes = [C.POINTER(_i64)] * 4
        L.amro_topology_cache_stats.restype = None
        L.amro_topology_cache_clear.restype = None
        L.amro_simulate_discrete_model.argtypes = [C.c_void_p, _i64, _i64, _i64] * 5 + [C.c_int] * 5 + [C.c_void_p]
        L.amro_simulate_and_collapse.argtypes = [C.c_void_p, _i64, _i64, _i64] * 2 + [C.c_double, C.c_double, _i64] + \
            [C.c_double] * 6 + [_i64] * 4 + [C.c_int, C.c_int, C.c_int, C.POINTER(_i64), C.c_void_p]
        L.amro_progress_patients.argtypes = [C.c_void_p, _i64, _i64, _i64, C.c_void_p, _i64, _i64, _i64, C.c_double,
                                             C.c_void_p, _i64, _i64, _i64, _i64, C.c_int, C.c_int, C.c_int, C.c_int,
                                             _i64, _i64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.amro_total_positive.argtypes = [C.c_void_p, _i64, _i64, _i64, _i64, C.c_int, C.c_int, C.POINTER(_i64),
                                          C.POINTER(_i64), C.c_void_p]
        L.amro_summary_of_total_positive.argtypes = [C.c_void_p, _i64, _i64, _i64, C.c_void_p, _i64, C.c_void_p]
        L.amro_summary_of_total_positive_device.argtypes = [C.c_void_p, _i64, _i64, _i64, C.c_void_p, _i64, C.c_int, C.c_void_p]
        L.amro_day_moments_peers.argtypes = [C.c_void_p, C.c_void_p, _i64, _i64, _i64, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_void_p]
        L.amro_topology_debug_arrays.argtypes = [C.c_void_p] + [C.POINTER(C.c_void_p)] * 11 + [C.POINTER(_i64)]
        _lib = L
    return _lib


def cache_stats() -> dict:
    """Hits / misses / entries / bytes of the compiled-hospital cache behind the one-call entry points."""
    v = [_i64() for _ in range(4)]
    lib().amro_topology_cache_stats(*[C.byref(x) for x in v])
    return dict(zip(("hits", "misses", "entries", "bytes"), (x.value for x in v)))


def check(rc: int):
    """Raise what the reference raises for this error class."""
    if rc == AMRO_OK:
        return
    msg = lib().amro_last_error().decode()
    if rc == AMRO_EINDEX:
        raise IndexError(msg)
    if rc == AMRO_EINVAL:
        raise ValueError(msg)
    raise RuntimeError(msg)
