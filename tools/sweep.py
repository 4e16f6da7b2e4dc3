"""Launch-shape sweeps of the fused kernels on one hospital (development tool; results go to profiles/sweeps.txt).

  python tools/sweep.py cfg2 "AMRO_BITS_G=9" "AMRO_BITS_G=12 AMRO_BITS_CTAS_PER_SM=3" ...

Every argument after the workload is one variant: space-separated NAME=VALUE pairs put into the environment before the
runs (the library reads them per call).  Prints the best kernel time of 5 runs per variant and a checksum of the per-day
counts (all variants of one seed must agree).
"""
import os
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from abm_amro_b200 import engine, synth  # noqa: E402

SHAPES = {"cfg2": (30, 10_000, 365, 1000), "cfg3": (20, 5_000, 180, 65_536), "cfg3s": (20, 5_000, 180, 16_384),
          "cfg3x8": (20, 5_000, 180, 8_192), "oneward": (1, 10_000, 60, 1000), "twoward64": (2, 20_000, 30, 64), "cfg4s": (500, 200_000, 30, 256), "mid": (30, 40_000, 60, 768)}


def main():
    name = sys.argv[1]
    nw, cen, nd, S = SHAPES[name]
    wm, tp, n_init = synth.make_hospital(nw, cen, nd, seed=1)
    par = synth.tutorial_parameters(S) if not name.startswith("cfg3") else synth.particle_parameters(S, seed=100)
    updates = float(wm.shape[0]) * S
    managed = set()
    with engine.Topology(wm, tp, n_init) as topo:
        for variant in sys.argv[2:] or [""]:
            for k in managed:
                os.environ.pop(k, None)
            for kv in variant.split():
                k, v = kv.split("=", 1)
                os.environ[k] = v
                managed.add(k)
            best, crc = 1e30, None
            for _ in range(5):
                r = topo.simulate(par, 0.2, 0.2, 2, 7, 1, 42)
                best = min(best, r.kernel_ms)
                crc = zlib.crc32(r.counts_colonized.tobytes()) ^ zlib.crc32(r.counts_detected.tobytes())
            print(f"{name} [{variant or 'default'}] kernel_ms {best:.3f}  {updates / best / 1e6:.1f} G updates/s  checksum {crc:08x}", flush=True)


if __name__ == "__main__":
    main()
