"""TEST INFRASTRUCTURE -- runs ONE call of the compiled, unmodified reference (oracle/_ref/<variant>/amro*.so)
in its own process.  Harness rules for this image (numpy 2 vs the reference's numpy<2 pin):
pass only float64 F-contiguous arrays, copy results and drop them before interpreter exit.

usage: python _ref_worker.py <native|replay> <in.npz> <out.npz>
in.npz: 'func' (str), 'argnames' (list), one array per positional argument, optional 'stream' (replay)
"""
import ctypes
import os
import sys
import time

import numpy as np

variant, fin, fout = sys.argv[1:4]
here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(here, "_ref", variant))
import amro  # noqa: E402  (the reference module)

z = np.load(fin, allow_pickle=False)
func = str(z["func"])
args = []
for name in [str(x) for x in z["argnames"]]:
    a = z[name]
    if a.ndim == 0:
        v = a.item()
        args.append(v)
    else:
        args.append(np.asfortranarray(a.astype(np.float64)))
pos = -1
if variant == "replay":
    lib = ctypes.CDLL(amro.__file__)
    lib.amro_replay_set.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.c_size_t]
    lib.amro_replay_pos.restype = ctypes.c_size_t
    stream = np.ascontiguousarray(z["stream"], dtype=np.float64)
    lib.amro_replay_set(stream.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), stream.size)
reps = int(z["reps"]) if "reps" in z else 1
t0 = time.perf_counter()
for _ in range(reps):
    res = getattr(amro, func)(*args)
dt = (time.perf_counter() - t0) / reps
out = np.array(res, copy=True)
del res
if variant == "replay":
    pos = lib.amro_replay_pos()
# in-place mutation of a borrowed input (progress_* functions) is part of the contract: return it too
extra = {}
if func.startswith("progress_"):
    extra["mutated_input"] = np.array(args[0], copy=True)
np.savez(fout, out=out, seconds=dt, stream_pos=pos, **extra)
