"""Where the time of a FIRST run() goes (one B200): page-locking the reader's array, allocation, staging, compute --
and what the same run costs from pageable memory."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from bench import synthetic_trajectory  # noqa: E402
from transport_analysis_b200 import _lib  # noqa: E402
from transport_analysis_b200.synthetic import make_universe  # noqa: E402
from transport_analysis_b200.velocityautocorr import VelocityAutocorr  # noqa: E402


def clock(label, fn):
    t0 = time.perf_counter()
    out = fn()
    print(f"{label:58s} {1e3 * (time.perf_counter() - t0):9.1f} ms", flush=True)
    return out


T, A = 10000, 100000
vel = synthetic_trajectory(T, A, seed=1)
vel2 = synthetic_trajectory(T, A, seed=2)
ctx = clock("context (CUDA initialisation)", lambda: _lib.Context([0]))
clock("page-lock 12 GB (first registration of the process)", lambda: _lib.pin_array(vel))
clock("page-lock another 12 GB", lambda: _lib.pin_array(vel2))
u = make_universe(None, vel)
clock("run(), array page-locked, new context", lambda: VelocityAutocorr(u.atoms, fft=True).run())
vel3 = synthetic_trajectory(T, A, seed=3)
u3 = make_universe(None, vel3)
clock("run(), pageable array, pin_host=False", lambda: VelocityAutocorr(u3.atoms, fft=True, pin_host=False).run())
vel4 = synthetic_trajectory(T, A, seed=4)
u4 = make_universe(None, vel4)
clock("run(), pageable array, default (the class page-locks it)", lambda: VelocityAutocorr(u4.atoms, fft=True).run())
a = VelocityAutocorr(u4.atoms, fft=True)
clock("  second analysis object on the same array", lambda: a.run())
clock("  second run() of that object", lambda: a.run())
