"""TEST INFRASTRUCTURE -- subprocess front-end of the compiled reference (see _ref_worker.py)."""
from __future__ import annotations

import glob
import os
import subprocess
import sys
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def available(variant: str = "native") -> bool:
    return bool(glob.glob(os.path.join(_HERE, "_ref", variant, "amro*.so")))


def build() -> bool:
    """Build oracle/_ref from /root/reference when it is present (never on the GPU box)."""
    if not os.path.isdir("/root/reference/src/amro"):
        return False
    subprocess.run(["make", "-C", _HERE, "ref"], check=True, stdout=subprocess.DEVNULL)
    return True


def run(func: str, args: list, argnames: list[str] | None = None, variant: str = "native", stream=None,
        reps: int = 1, timeout: float = 3600.0) -> dict:
    """Call ``amro.<func>(*args)`` of the compiled reference; returns {'out', 'seconds', 'stream_pos', ...}."""
    argnames = argnames or [f"a{i}" for i in range(len(args))]
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.npz"), os.path.join(td, "out.npz")
        payload = {n: np.asarray(a) for n, a in zip(argnames, args)}
        payload.update(func=np.array(func), argnames=np.array(argnames), reps=np.array(reps))
        if stream is not None:
            payload["stream"] = np.asarray(stream, dtype=np.float64)
        np.savez(fin, **payload)
        p = subprocess.run([sys.executable, os.path.join(_HERE, "_ref_worker.py"), variant, fin, fout],
                           capture_output=True, text=True, timeout=timeout)
        if p.returncode != 0:
            raise RuntimeError(f"reference worker failed ({p.returncode}): {p.stderr[-2000:]}")
        z = np.load(fout)
        return {k: z[k] for k in z.files}
