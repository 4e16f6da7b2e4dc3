"""Handle-based Python front-end of the C ABI: compile a hospital once, simulate it many times.

``Topology`` wraps ``amro_topology_create``; ``Topology.simulate`` wraps ``amro_simulate`` (host numpy
buffers, copies inside the call) and ``Topology.simulate_device`` the same entry point with device
pointers (any object exposing ``data_ptr()``, e.g. a torch CUDA tensor).  The drop-in functions of the
reference live in the compiled module ``abm_amro_b200.amro``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import capi


def _f64(a):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    if a.ndim != 2:
        raise RuntimeError("CARMA: Number of dimensions must be 1 <= ndim <= 2")
    return np.asfortranarray(a)


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


@dataclass
class SimResult:
    counts_colonized: np.ndarray | None   # [n_days_out x S] int32
    counts_detected: np.ndarray | None
    state: np.ndarray | None              # [R x 2S] float64: C columns then D columns
    p: np.ndarray | None                  # [R x S] transition probabilities (replay / general path)
    pressure: np.ndarray | None           # [n_segments x S] int32 (fast path)
    moments: np.ndarray | None            # [n_days_out x 4]: sum, sum^2 of colonised counts, sum, sum^2 of detected counts
    path_used: int
    launches: int
    kernel_ms: float
    total_ms: float
    distance: np.ndarray | None = None    # [S] squared distance of the per-day counts to observed series, if asked for
    kernel_kind: int = 0                  # capi.KERNEL_*: which kernel stepped the records
    grid: int = 0                         # CTAs of the fused kernel's launch
    ctas_per_team: int = 0
    team_barrier: int = 0                 # capi.BARRIER_*
    summary_colonized: np.ndarray | None = None   # [n_days_out x (3 + nq)]: day, mean, sd, quantiles over the simulations
    summary_detected: np.ndarray | None = None


def hospital_spec(n_wards: int, census: int, n_days: int, seed: int = 1, p_discharge: float = 0.15, p_transfer: float = 0.10):
    """The synthetic "bed model" hospital of include/amro_b200.h (amro_hospital_spec)."""
    return capi.HospitalSpec(int(n_wards), int(census), int(n_days), int(seed), float(p_discharge), float(p_transfer))


def hospital_tables(spec):
    """The hospital of ``spec`` as the reference's float64 tables (small sizes): ``(ward_matrix, tppw, n_init)``."""
    R, K = capi._i64(), capi._i64()
    capi.check(capi.lib().amro_hospital_tables(C.byref(spec), C.byref(R), C.byref(K), None, None))
    wm = np.zeros((R.value, 6), order="F"); tp = np.zeros((K.value, 3), order="F")
    capi.check(capi.lib().amro_hospital_tables(C.byref(spec), C.byref(R), C.byref(K), _ptr(wm), _ptr(tp)))
    return wm, tp, int(spec.census)


class Topology:
    """A hospital (ward_matrix + total_patients_per_ward) compiled for one GPU."""

    def __init__(self, ward_matrix, total_patients_per_ward, n_init: int, device: int = 0, _handle=None):
        self._h = C.c_void_p()
        if _handle is not None:
            self._h = _handle
        else:
            wm, tp = _f64(ward_matrix), _f64(total_patients_per_ward)
            capi.check(capi.lib().amro_topology_create(_ptr(wm), wm.shape[0], wm.shape[1], wm.shape[0],
                                                       _ptr(tp), tp.shape[0], tp.shape[1], tp.shape[0],
                                                       int(n_init), int(device), C.byref(self._h)))
        info = capi.TopologyInfo()
        capi.check(capi.lib().amro_topology_get_info(self._h, C.byref(info)))
        self.info = info
        self.device = device

    @classmethod
    def generate(cls, spec, device: int = 0) -> "Topology":
        """A synthetic hospital generated on the device straight into the compiled form (no host tables; fused kernels only)."""
        h = C.c_void_p()
        capi.check(capi.lib().amro_topology_generate(C.byref(spec), int(device), C.byref(h)))
        return cls(None, None, 0, device, _handle=h)

    def device_arrays(self) -> dict:
        """The compiled form as it sits on the device (tests: the generator against the host compiler)."""
        i = self.info
        D = i.total_days + 1
        out = {"meta": np.zeros((i.n_rows_stepped, 4), dtype=np.uint32), "seg_row_start": np.zeros(i.n_segments + 1, dtype=np.int64),
               "seg_N": np.zeros(i.n_segments), "seg_main_next": np.zeros(i.n_segments, dtype=np.uint32),
               "day_seg_start": np.zeros(D + 1, dtype=np.int64), "init_seg": np.zeros(i.n_init, dtype=np.uint32),
               "init_slot": np.zeros(i.n_init, dtype=np.uint32)}
        capi.check(capi.lib().amro_topology_debug_device_arrays(self._h, *[C.c_void_p(v.ctypes.data) for v in out.values()]))
        return out

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            capi.lib().amro_topology_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def kernel_times(self, n: int) -> np.ndarray:
        """Device time (ms) of the fused kernel of the last ``n`` runs, oldest first; waits for them (asynchronous runs)."""
        ms = (C.c_float * n)()
        capi.check(capi.lib().amro_topology_kernel_times(self._h, n, ms))
        return np.array(ms[:], dtype=np.float64)

    # ---- introspection (host arrays of the compiled form; used by the CPU tests of the compiler) ----
    def debug_arrays(self) -> dict:
        names = ["order", "day_start", "day_seg_start", "seg_row_start", "seg_N", "meta_next", "meta_info",
                 "meta_nseg", "init_seg", "push_src", "push_dst"]
        ptrs = [C.c_void_p() for _ in names]
        n_push = C.c_int64()
        capi.check(capi.lib().amro_topology_debug_arrays(self._h, *[C.byref(p) for p in ptrs], C.byref(n_push)))
        i = self.info
        D = i.total_days + 1
        shapes = {"order": (i.n_rows, np.int64), "day_start": (D + 1, np.int64), "day_seg_start": (D + 1, np.int64),
                  "seg_row_start": (i.n_segments + 1, np.int64), "seg_N": (i.n_segments, np.float64),
                  "meta_next": (i.n_rows_stepped, np.uint32), "meta_info": (i.n_rows_stepped, np.uint32),
                  "meta_nseg": (i.n_rows_stepped, np.uint32), "init_seg": (i.n_init, np.uint32),
                  "push_src": (n_push.value, np.int64), "push_dst": (n_push.value, np.int64)}
        out = {}
        for name, p in zip(names, ptrs):
            n, dt = shapes[name]
            if not p.value or n == 0:
                out[name] = np.zeros(0, dtype=dt) if p.value or n == 0 else None
                continue
            buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(p.value)
            out[name] = np.frombuffer(buf, dtype=dt).copy()
        return out

    # ---- host-buffer run (what a user of the reference calls; copies are inside the call) -------------
    def simulate(self, parameters, icp, idp, time_to_detect: int, tsh: int, tsa: int, seed: int, *,
                 n_sims: int | None = None, sim_offset: int = 0, rng_mode: int = capi.RNG_PHILOX,
                 path: int = capi.PATH_AUTO, keep_trajectory: bool = False, want_counts: bool = True,
                 want_state: bool = False, want_p: bool = False, want_pressure: bool = False,
                 want_moments: bool = False, u_col=None, u_det=None, u_icol=None, u_idet=None,
                 observed_colonized=None, observed_detected=None, quantiles=None) -> SimResult:
        """``observed_*``: per-day series (NaN = no data); the result then carries ``distance`` [S], the per-simulation sum
        of squared differences between the simulated per-day counts and the data, computed on the device."""
        i = self.info
        par = _f64(parameters)
        S = int(n_sims if n_sims is not None else par.shape[0])
        a = capi.SimArgs()
        a.n_sims, a.sim_offset = S, sim_offset
        a.time_to_detect, a.testing_schedule_hospitalized, a.testing_schedule_arrivals, a.seed = time_to_detect, tsh, tsa, seed
        a.rng_mode, a.path, a.keep_trajectory, a.pointers_on_device = rng_mode, path, int(keep_trajectory), 0
        a.parameters, a.p_rows, a.p_ld = par.ctypes.data, par.shape[0], par.shape[0]
        keep = [par]
        for name, val in (("icp", icp), ("idp", idp)):
            if np.isscalar(val):
                arr = np.array([float(val)])
                rs = cs = 0
            else:
                arr = _f64(val)
                if arr.shape != (i.n_init, S):
                    raise RuntimeError(f"{name}: expected shape ({i.n_init}, {S}), got {arr.shape}")
                rs, cs = 1, arr.shape[0]
            keep.append(arr)
            setattr(a, name, arr.ctypes.data); setattr(a, name + "_rs", rs); setattr(a, name + "_cs", cs)
        if rng_mode == capi.RNG_REPLAY:
            uc, ud, uic, uid = _f64(u_col), _f64(u_det), _f64(u_icol), _f64(u_idet)
            keep += [uc, ud, uic, uid]
            a.u_col, a.u_det, a.u_ld = uc.ctypes.data, ud.ctypes.data, uc.shape[0]
            a.u_icol, a.u_idet, a.ui_ld = uic.ctypes.data, uid.ctypes.data, uic.shape[0]
        cc = cd = st = p = pr = mo = None
        if want_counts:
            cc = np.zeros((i.n_days_out, S), dtype=np.int32, order="F"); cd = np.zeros_like(cc)
            a.counts_colonized, a.counts_detected = cc.ctypes.data, cd.ctypes.data
        if want_state:
            st = np.empty((i.n_rows, 2 * S), order="F"); a.state_out, a.so_ld = st.ctypes.data, i.n_rows
        if want_p:
            p = np.empty((i.n_rows, S), order="F"); a.p_out = p.ctypes.data
        if want_pressure:
            pr = np.zeros((i.n_segments, S), dtype=np.int32, order="F"); a.pressure_out = pr.ctypes.data
        if want_moments:
            mo = np.zeros((i.n_days_out, 4), order="F"); a.day_moments = mo.ctypes.data
        dist = None
        if observed_colonized is not None or observed_detected is not None:
            dist = np.zeros(S)
            a.distance = dist.ctypes.data
            for name, val in (("observed_colonized", observed_colonized), ("observed_detected", observed_detected)):
                if val is not None:
                    arr = np.ascontiguousarray(val, dtype=np.float64)
                    if arr.shape != (i.n_days_out,):
                        raise RuntimeError(f"{name}: expected shape ({i.n_days_out},), got {arr.shape}")
                    keep.append(arr)
                    setattr(a, name, arr.ctypes.data)
        sc = sd = None
        if quantiles is not None:   # summary_of_total_positive of the run's counts, computed on the device
            qs = np.ascontiguousarray(quantiles, dtype=np.float64).ravel()
            sc = np.zeros((i.n_days_out, 3 + qs.size), order="F"); sd = np.zeros_like(sc)
            keep.append(qs)
            a.quantiles, a.n_quantiles = qs.ctypes.data, qs.size
            a.summary_colonized, a.summary_detected = sc.ctypes.data, sd.ctypes.data
        capi.check(capi.lib().amro_simulate(self._h, C.byref(a)))
        r = SimResult(cc, cd, st, p, pr, mo, a.path_used, a.launches, a.kernel_ms, a.total_ms)
        r.distance = dist
        r.summary_colonized, r.summary_detected = sc, sd
        r.kernel_kind, r.grid, r.ctas_per_team, r.team_barrier = a.kernel_kind, a.grid, a.ctas_per_team, a.team_barrier
        return r

    # ---- device-resident run (inputs already in HBM; nothing crosses PCIe) ---------------------------------
    def simulate_device(self, parameters_ptr: int, p_rows: int, p_ld: int, icp_ptr: int, idp_ptr: int,
                        counts_col_ptr: int, counts_det_ptr: int, n_sims: int, time_to_detect: int, tsh: int,
                        tsa: int, seed: int, *, sim_offset: int = 0, stream: int = 0,
                        icp_strides=(0, 0), idp_strides=(0, 0), path: int = capi.PATH_AUTO,
                        keep_trajectory: bool = False, day_moments_ptr: int = 0, asynchronous: bool = False) -> SimResult:
        """``asynchronous``: enqueue on ``stream`` and return at once (no timings); the caller synchronises the stream."""
        a = capi.SimArgs()
        a.n_sims, a.sim_offset = n_sims, sim_offset
        a.time_to_detect, a.testing_schedule_hospitalized, a.testing_schedule_arrivals, a.seed = time_to_detect, tsh, tsa, seed
        a.rng_mode, a.path, a.keep_trajectory, a.pointers_on_device = capi.RNG_PHILOX, path, int(keep_trajectory), 1
        a.parameters, a.p_rows, a.p_ld = parameters_ptr, p_rows, p_ld
        a.icp, a.icp_rs, a.icp_cs = icp_ptr, icp_strides[0], icp_strides[1]
        a.idp, a.idp_rs, a.idp_cs = idp_ptr, idp_strides[0], idp_strides[1]
        a.counts_colonized, a.counts_detected = counts_col_ptr, counts_det_ptr
        a.day_moments = day_moments_ptr or None
        a.stream = stream
        a.async_ = int(asynchronous)
        capi.check(capi.lib().amro_simulate(self._h, C.byref(a)))
        r = SimResult(None, None, None, None, None, None, a.path_used, a.launches, a.kernel_ms, a.total_ms)
        r.kernel_kind, r.grid, r.ctas_per_team, r.team_barrier = a.kernel_kind, a.grid, a.ctas_per_team, a.team_barrier
        return r
