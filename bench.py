#!/usr/bin/env python
"""Benchmark of the ward-level colonisation path (BASELINE.json metric: patient-day-updates/sec).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is ONE full run of the hot path over the workload: every day of the synthetic hospital for every
ensemble member resident on the GPU (config 2 of BASELINE.json: 30 wards x 10,000 patients x 365 days,
1,000 members per GPU; weak scaling: every GPU runs its own 1,000 members, Philox counters offset by the
global member index; --scaling strong splits the workload's members over the GPUs instead).  One JSON line is
printed by rank 0.

  value        device-resident: parameters already in HBM, per-day counts stay in HBM; the runs are enqueued
               asynchronously, back to back, with the run's one collective (N > 1) on the same stream (CUDA events)
  e2e          the same run through the host-pointer C ABI call a user makes (amro_simulate with pinned host
               buffers): parameters copied host->device and the per-(day, member) counts copied back, every step
  roofline     algorithmic bytes (4 + 20/S_gpu per update) / device time of the fused kernel
  cpu_baseline the reference's own CPU implementation (oracle/_ref, compiled from the unmodified sources) on the
               host cores, on a bounded sample of the same workload
  e2e_dropin   the reference's own entry point (amro.simulate_and_collapse of the drop-in module), numpy in / out
  secondary    config 3 (the 64k-particle sweep) timed in the same process, so that its numbers are driver-run too
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (wards, census, days, members per GPU)
    "cfg1_5w_300p_100d_20s": (5, 300, 100, 20),
    "cfg2_30w_10kp_365d_1000s": (30, 10_000, 365, 1000),
    "cfg3_20w_5kp_180d_65536s": (20, 5_000, 180, 65_536),
    # regional network at its full size: 3.65e8 patient-day rows, and the national network: 7.3e9 rows, 32 members per GPU.
    # Their hospitals are generated ON THE DEVICE straight into the compiled form (engine.Topology.generate: the tables never
    # exist on the host); not default bench lines, run with --workload and a few steps
    "cfg4_500w_1Mp_365d_256s": (500, 1_000_000, 365, 256),
    "cfg5_5kw_10Mp_730d_32s": (5_000, 10_000_000, 730, 32),
}
GENERATED = {"cfg4_500w_1Mp_365d_256s", "cfg5_5kw_10Mp_730d_32s"}
DEFAULT_WORKLOAD = "cfg2_30w_10kp_365d_1000s"
TTD, TSH, TSA, SEED, ICP, IDP = 2, 7, 1, 42, 0.2, 0.2


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_bytes(workload: str):
    """DRAM bytes per launch of the fused kernel from the committed ncu --set full capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(workload)
        except Exception:
            return None
    return None


# ----------------------------------------------------------------------------------------------------------
# clocks during the timed region
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int, period_s: float = 0.02):
        self.index, self.period, self.samples, self.reasons, self.max_mhz = index, period_s, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.nv:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation on the host cores, bounded sample
# ----------------------------------------------------------------------------------------------------------
def truncate_hospital(wm, tp, n_days):
    keep = wm[:, 0] < n_days
    w = np.asfortranarray(wm[keep].copy())
    w[w[:, 5] >= w.shape[0], 5] = -1.0  # links into the removed days
    t = np.asfortranarray(tp[tp[:, 0] < n_days].copy())
    return w, t


def cpu_sample(workload: str, sample_days: int = 60, sims_per_proc: int = 32, reps: int = 1):
    """Run the reference CPU path on `procs` host processes at once, each stepping `sims_per_proc` ensemble members of the
    first `sample_days` days of the workload (every process gets the same inputs and seed: the reference's cost does not
    depend on them, and its output is discarded).  Returns (updates/s, procs, kind, text)."""
    from abm_amro_b200 import synth
    from oracle import ref_runner
    wards, census, days, _ = WORKLOADS[workload]
    sample_days = max(1, min(sample_days, 600_000 // census))   # about 6e5 patient-day rows per process
    wm, tp, n_init = synth.make_hospital(wards, census, min(days, sample_days), seed=1)
    R = wm.shape[0]
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        avail = int(open("/proc/meminfo").read().split("MemAvailable:")[1].split()[0]) * 1024
    except Exception:
        avail = 16 << 30
    per_proc = R * (6 + 2 * sims_per_proc) * 8 * 3
    procs = max(1, min(ncpu, int(avail * 0.5 / per_proc)))
    par = synth.tutorial_parameters(sims_per_proc)
    icp = np.full((n_init, sims_per_proc), ICP, order="F")
    idp = np.full((n_init, sims_per_proc), IDP, order="F")
    updates = float(R) * sims_per_proc * procs
    if ref_runner.available("native"):
        kind = "reference"
        import tempfile
        with tempfile.TemporaryDirectory() as td:
            fin = os.path.join(td, "in.npz")
            names = ["icp", "idp", "wm", "tp", "par", "ttd", "tsh", "tsa", "seed"]
            vals = [icp, idp, wm, tp, par, np.array(TTD), np.array(TSH), np.array(TSA), np.array(SEED)]
            np.savez(fin, func=np.array("simulate_discrete_model"), argnames=np.array(names), reps=np.array(reps),
                     **dict(zip(names, vals)))
            worker = os.path.join(ROOT, "oracle", "_ref_worker.py")
            ps = [subprocess.Popen([sys.executable, worker, "native", fin, os.path.join(td, f"o{i}.npz")],
                                   stdout=subprocess.DEVNULL, stderr=subprocess.PIPE) for i in range(procs)]
            for p in ps:
                _, err = p.communicate()
                if p.returncode != 0:
                    raise RuntimeError("reference worker failed: " + err.decode()[-500:])
            secs = max(float(np.load(os.path.join(td, f"o{i}.npz"))["seconds"]) for i in range(procs))
    else:
        kind = "port"
        import multiprocessing as mp
        with mp.get_context("spawn").Pool(procs) as pool:
            secs = max(pool.map(_port_worker, [(wards, census, min(days, sample_days), sims_per_proc, reps)] * procs))
    text = (f"{workload}: first {min(days, sample_days)} days ({R} patient-day rows) x {sims_per_proc} members per process, "
            f"{procs} concurrent single-threaded processes (the reference has no threading), simulate_discrete_model only")
    return updates / secs, procs, kind, text


def _port_worker(spec):
    wards, census, days, S, reps = spec
    sys.path.insert(0, ROOT)
    from abm_amro_b200 import synth
    from oracle import oracle
    wm, tp, n_init = synth.make_hospital(wards, census, days, seed=1)
    par = synth.tutorial_parameters(S)
    icp = np.full((n_init, S), ICP, order="F"); idp = np.full((n_init, S), IDP, order="F")
    t0 = time.perf_counter()
    for _ in range(reps):
        oracle.simulate_discrete_model(icp, idp, wm, tp, par, TTD, TSH, TSA, SEED)
    return (time.perf_counter() - t0) / reps


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    ts = []
    val = procs = kind = text = None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        val, procs, kind, text = cpu_sample(args.workload)
        if i >= args.warmup:
            ts.append((time.perf_counter() - t0, val))
    value = float(np.mean([v for _, v in ts]))
    wards, census, days, S = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": "patient-day-updates/sec", "value": value, "unit": "updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": float(np.mean([t for t, _ in ts])) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "wards": wards, "census": census, "days": days, "members_per_gpu": S,
                   "note": "CPU arm: bounded sample of the workload, see cpu_baseline.sample"},
        "cpu_baseline": {"value": value, "unit": "updates/s", "cores": procs, "kind": kind, "sample": text},
        "e2e": {"value": value, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------
def _algorithmic_bytes(S):
    return 4.0 + 20.0 / S        # 2 B state read + 2 B state write + the record's topology shared by S members


def time_workload(topo, par_np, S, sim_offset, dev, steps, warmup, world, dist, flush, local_rank):
    """`steps` device-resident runs of one compiled hospital, enqueued back to back on the current stream (asynchronous C ABI
    calls: kernel -> the run's collective (N > 1: peer-store moments kernel, or day moments + NCCL all-reduce), no host synchronisation
    inside the timed region).  Returns a dict."""
    import torch
    from abm_amro_b200 import ensemble
    n_days = topo.info.n_days_out
    par_d = torch.from_numpy(np.ascontiguousarray(par_np.T)).to(dev)            # [6 x S] row-major == [S x 6] column-major
    icp_d = torch.tensor([ICP], dtype=torch.float64, device=dev)
    idp_d = torch.tensor([IDP], dtype=torch.float64, device=dev)
    cc_d = torch.zeros((S, n_days), dtype=torch.int32, device=dev)               # == [n_days x S] column-major
    cd_d = torch.zeros((S, n_days), dtype=torch.int32, device=dev)
    # The run's collective (N > 1): the per-day moments of every run, delivered to all ranks.  Peer stores when the ranks can map
    # each other's memory (ensemble.PeerMoments: one small kernel per run, no rank waits for another), else an NCCL all-reduce per
    # run on NCCL's stream (two buffers, so that the reduction of run k overlaps run k + 1).
    mom = [torch.zeros((4, n_days), dtype=torch.float64, device=dev) for _ in range(2)]   # == [n_days x 4] column-major
    pending = [None, None]
    stream = torch.cuda.current_stream(dev)
    count = [0]
    peers = None
    no_coll = bool(os.environ.get("AMRO_BENCH_NO_COLLECTIVE"))          # (development switch: what the collective costs)
    if world > 1 and not no_coll and not os.environ.get("AMRO_BENCH_NCCL"):
        try:
            peers = ensemble.PeerMoments(n_days, dev)
        except Exception as e:                                            # noqa: BLE001 -- any failure of the rendezvous: NCCL does it
            print(f"[bench] peer stores unavailable ({type(e).__name__}: {e}); the collective is an NCCL all-reduce", file=sys.stderr)
            peers = None

    def step_device():
        k = count[0] & 1
        count[0] += 1
        flush.zero_()                                                            # evict L2 between steps
        use_nccl = world > 1 and peers is None and not no_coll
        if pending[k] is not None:
            pending[k].wait()                                                    # (stream-side wait) run k - 2's reduction has read this buffer
        r = topo.simulate_device(par_d.data_ptr(), S, S, icp_d.data_ptr(), idp_d.data_ptr(), cc_d.data_ptr(),
                                 cd_d.data_ptr(), S, TTD, TSH, TSA, SEED, sim_offset=sim_offset, stream=stream.cuda_stream,
                                 day_moments_ptr=mom[k].data_ptr() if use_nccl else 0, asynchronous=True)
        if peers is not None:
            peers.publish(count[0] - 1, cc_d.data_ptr(), cd_d.data_ptr(), S, n_days, stream=stream.cuda_stream)
            r.launches += 1
        elif use_nccl:  # per-day (sum, sum^2) of both metrics over all members, summed in place
            pending[k] = dist.all_reduce(mom[k], op=dist.ReduceOp.SUM, async_op=True)
        return r

    def finish():
        for w in pending:
            if w is not None:
                w.wait()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(warmup):
        step_device()
    finish()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    with ClockSampler(local_rank) as clocks:
        e0.record(stream)
        for _ in range(steps):
            r = step_device()
            launches += r.launches
        finish()                                                                 # the last reductions, inside the timed region
        e1.record(stream)
        barrier()
    elapsed_ms = e0.elapsed_time(e1)
    collective = "none (1 GPU)" if world == 1 else ("skipped (development switch)" if no_coll else "nccl all_reduce per run")
    mom_last = mom[(count[0] - 1) & 1]
    if peers is not None:
        # outside the timed region: the last run's moments as every rank received them == an NCCL all-reduce of the same run's
        got = peers.reduced(count[0] - 1)
        want = ensemble.day_moments(cc_d, cd_d).T.contiguous()
        dist.all_reduce(want, op=dist.ReduceOp.SUM)
        if not torch.equal(got, want):
            raise SystemExit("peer-store collective disagrees with the all-reduce of the same moments")
        collective = "peer stores over NVLink from the moments kernel (torch symmetric memory); checked against an NCCL all-reduce of the last run"
        mom_last = got
    kernel_ms = topo.kernel_times(min(steps, 64))                               # CUDA events around each run's fused kernel
    kernel_ms_ranks = None
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
        # every rank's mean kernel time (outside the timed region): the step is the slowest GPU's
        g = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(g, torch.tensor([float(np.mean(kernel_ms))], dtype=torch.float64, device=dev))
        kernel_ms_ranks = [round(float(x.item()), 4) for x in g]
    return dict(elapsed_ms=elapsed_ms, kernel_ms=float(np.mean(kernel_ms)), kernel_ms_ranks=kernel_ms_ranks, collective=collective, launches=launches, clocks=clocks.summary(), result=r,
                cc_d=cc_d, cd_d=cd_d, mom_d=mom_last, par_d=par_d)


def run_ours(args, rank, world, local_rank):
    import ctypes
    import torch
    from abm_amro_b200 import capi, engine, ensemble, synth

    wards, census, days, S_total = WORKLOADS[args.workload]
    if args.members:
        S_total = args.members
    strong = args.scaling == "strong"
    if strong:                       # the workload's members are split over the GPUs (the 64k-particle sweep: 8,192 per GPU on 8)
        sim_offset, S = ensemble.shard_sims(S_total, world, rank)
        if S == 0:
            raise SystemExit("strong scaling: more GPUs than groups of 32 members")
    else:                            # every GPU runs its own S_total members
        S = S_total
        sim_offset = rank * ((S + 63) // 64 * 64)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    out = sys.stdout
    if world > 1:
        # libraries write to file descriptor 1 (NCCL prints its version there): keep the real stdout for the ONE JSON line
        out = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    sweep = args.workload.startswith("cfg3")
    generated = args.workload in GENERATED
    if args.days:
        days = args.days
    torch.zeros(1, device=dev)                                                   # the CUDA context exists before the compile is timed
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    if generated:
        topo = engine.Topology.generate(engine.hospital_spec(wards, census, days, seed=1), device=local_rank)
        wm = tp = None
        n_init = census
        R = topo.info.n_rows
    else:
        wm, tp, n_init = synth.make_hospital(wards, census, days, seed=1)
        R = wm.shape[0]
        t0 = time.perf_counter()
        topo = engine.Topology(wm, tp, n_init, device=local_rank)
    topo_ms = (time.perf_counter() - t0) * 1e3
    n_days = topo.info.n_days_out
    total_members = S_total if strong else S * world
    updates_per_step = float(R) * total_members                                  # whole job, all GPUs
    par_all = synth.particle_parameters(S_total) if sweep else synth.tutorial_parameters(S_total)
    par_np = par_all[sim_offset:sim_offset + S] if strong else par_all
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)                # > 126 MB L2

    m = time_workload(topo, par_np, S, sim_offset, dev, args.steps, args.warmup, world, dist, flush, local_rank)
    value = updates_per_step * args.steps / (m["elapsed_ms"] * 1e-3)
    r = m["result"]
    if world > 1 and not os.environ.get("AMRO_BENCH_NO_COLLECTIVE"):  # what simulate_and_collapse reports, from the last step's reduced moments: [n_days x 4]
        stats = ensemble.moments_to_collapse(m["mom_d"].t(), total_members)
        assert bool(torch.isfinite(stats).all()) and float(stats[:, 0].max()) > 0.0

    # ---- end to end through the host-pointer C ABI call (pinned host buffers), the run's collective included -------------
    par_h = torch.empty((6, S), dtype=torch.float64).pin_memory()
    par_h.copy_(torch.from_numpy(np.ascontiguousarray(par_np.T)))
    cc_h = torch.empty((S, n_days), dtype=torch.int32).pin_memory()
    cd_h = torch.empty((S, n_days), dtype=torch.int32).pin_memory()
    mom_h = torch.empty((4, n_days), dtype=torch.float64).pin_memory()
    mom_e = torch.empty((4, n_days), dtype=torch.float64, device=dev)
    a = capi.SimArgs()
    a.n_sims, a.sim_offset = S, sim_offset
    a.time_to_detect, a.testing_schedule_hospitalized, a.testing_schedule_arrivals, a.seed = TTD, TSH, TSA, SEED
    a.rng_mode, a.path, a.keep_trajectory, a.pointers_on_device = capi.RNG_PHILOX, capi.PATH_AUTO, 0, 0
    a.parameters, a.p_rows, a.p_ld = par_h.data_ptr(), S, S
    icp_h, idp_h = np.array([ICP]), np.array([IDP])
    a.icp, a.idp = icp_h.ctypes.data, idp_h.ctypes.data
    a.counts_colonized, a.counts_detected = cc_h.data_ptr(), cd_h.data_ptr()
    if world > 1:
        a.day_moments = mom_h.data_ptr()

    def step_host():
        flush.zero_()
        capi.check(capi.lib().amro_simulate(topo._h, ctypes.byref(a)))
        if world > 1:   # the same collective as the device-resident leg: this rank's day moments, summed over the ranks
            mom_e.copy_(mom_h, non_blocking=True)
            ensemble.reduce_moments(mom_e)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(2):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = updates_per_step * e2e_steps / e2e_s
    same = bool(torch.equal(cc_h, m["cc_d"].cpu()) and torch.equal(cd_h, m["cd_d"].cpu()))

    # ---- the reference's own entry point (drop-in module, numpy in / numpy out, compiled-hospital cache warm) -------------
    dropin = None
    if world == 1 and not sweep and not generated and not args.no_dropin:
        from abm_amro_b200 import amro
        call = lambda: amro.simulate_and_collapse(wm, tp, ICP, IDP, n_init, .15, .55, .05, .6, .35, .9, S, TTD, TSH, TSA, SEED)  # noqa: E731
        call()
        t0 = time.perf_counter()
        for _ in range(5):
            call()
        dt = (time.perf_counter() - t0) / 5
        dropin = {"value": float(R) * S / dt, "unit": "updates/s", "ms_per_call": dt * 1e3,
                  "entry_point": "amro.simulate_and_collapse (compiled-hospital cache warm; the content hash of the tables is inside: it runs on host threads while the kernel runs on the hospital the quick key points to)"}

    # ---- the 64k-particle sweep in the same process (config 3), so that its fraction is driver-timed too --------------------
    secondary = None
    if world == 1 and args.workload == DEFAULT_WORKLOAD and not args.no_secondary:
        w3, c3, d3, S3 = WORKLOADS["cfg3_20w_5kp_180d_65536s"]
        wm3, tp3, n3 = synth.make_hospital(w3, c3, d3, seed=1)
        with engine.Topology(wm3, tp3, n3, device=local_rank) as topo3:
            m3 = time_workload(topo3, synth.particle_parameters(S3), S3, 0, dev, 5, 3, 1, None, flush, local_rank)
        u3 = float(wm3.shape[0]) * S3
        peak, _ = measured_peak_gbs()
        secondary = {"cfg3_20w_5kp_180d_65536s": {
            "value": u3 * 5 / (m3["elapsed_ms"] * 1e-3), "unit": "updates/s", "ms_per_step": m3["elapsed_ms"] / 5, "steps": 5, "warmup": 3,
            "kernel_ms": m3["kernel_ms"], "roofline_frac": u3 * _algorithmic_bytes(S3) / (m3["kernel_ms"] * 1e-3) / 1e9 / peak,
            "traffic": ncu_traffic_bytes("cfg3_20w_5kp_180d_65536s"), "clocks": m3["clocks"],
            "launch": {"grid": m3["result"].grid, "ctas_per_team": m3["result"].ctas_per_team}}}

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        b_alg = _algorithmic_bytes(S)
        k_ms = m["kernel_ms"]
        achieved = float(R) * S * b_alg / (k_ms * 1e-3) / 1e9                    # this GPU's kernel
        traffic = ncu_traffic_bytes(args.workload)
        cpu = None
        if not args.no_cpu_baseline and world == 1:   # the CPU baseline is timed at N = 1 only
            v, procs, kind, text = cpu_sample(args.workload)
            cpu = {"value": v, "unit": "updates/s", "cores": procs, "kind": kind, "sample": text}
        kinds = {capi.KERNEL_CODES: "16-bit state codes", capi.KERNEL_BITS: "bit-sliced state", capi.KERNEL_FLOAT64: "float64 planes"}
        line = {
            "metric": "patient-day-updates/sec", "value": value, "unit": "updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["elapsed_ms"] / args.steps,
            "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "u32 bit planes (32 members per register) / Philox4x32 / f64 probabilities",
            "data": "synthetic",
            "config": {"workload": args.workload, "wards": wards, "census": census, "days": days, "members_per_gpu": S,
                       "members_total": total_members, "patient_day_rows": R, "updates_per_step": updates_per_step,
                       "path": "fast" if r.path_used == capi.PATH_FAST else "general",
                       "kernel": kinds.get(r.kernel_kind, "?"),
                       "launch": {"grid": r.grid, "ctas_per_team": r.ctas_per_team, "team_barrier": ["cta", "global", "cluster"][r.team_barrier]},
                       "l2": "256 MB buffer written between steps (inside the timed region)",
                       "timed_region": "asynchronous runs enqueued back to back: kernel -> day moments -> the run's collective (N > 1), one sync at the end",
                       "collective": m["collective"],
                       "topology_compile_ms": topo_ms,
                       "hospital": "generated on the device (amro_topology_generate)" if generated else "numpy tables through the C ABI (synth.make_hospital)",
                       "counts_match_between_device_and_host_api": same},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_update": b_alg,
                         "kernel_ms": k_ms,
                         "note": "algorithmic bytes per SURVEY.md 8(d); the kernel itself keeps 4 bits per (record, member): it moves about "
                                 "1 B per update and is bound by the integer ALU pipe and, on config 2, by per-day latency -- not by HBM (DESIGN.md 3.5)",
                         "kernel_ms_ranks": m.get("kernel_ms_ranks")},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "updates/s",
                    "h2d_bytes_per_step": int(S * 6 * 8 + 16 + (4 * n_days * 8 if world > 1 else 0)),
                    "d2h_bytes_per_step": int(2 * n_days * S * 4 + (4 * n_days * 8 if world > 1 else 0)),
                    "steps": e2e_steps, "collective_inside": world > 1},
            "e2e_dropin": dropin,
            "secondary": secondary,
            "gpu_launches": m["launches"],
            "clocks": m["clocks"],
        }
        print(json.dumps(line), file=out, flush=True)
    topo.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--members", type=int, default=0, help="override members per GPU")
    ap.add_argument("--days", type=int, default=0, help="override the number of days (generated workloads: what fits in HBM)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the config-3 block of the default line")
    ap.add_argument("--no-dropin", action="store_true", help="skip timing the drop-in module's simulate_and_collapse")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every GPU runs the workload's members; strong: the members are split over the GPUs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        args.steps = args.steps if args.steps is not None else 3
        args.warmup = args.warmup if args.warmup is not None else 1
        run_reference_arm(args, rank, world)
        return
    args.steps = args.steps if args.steps is not None else 50
    args.warmup = args.warmup if args.warmup is not None else 5
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-launch ourselves the way the driver does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
