"""Measures the error constant C of the un-refined FFT route of the Helfand MSD: |S1 - 2 S2 - exact| <= C eps sum g^2,
per (particle, lag), on the trajectory families of tests/test_gpu_parity.py::_helfand_case.  The refinement threshold in
ta_helfand_fft assumes C = 100 and a target of 2e-11 (TA_B200_HELFAND_FFT_THR = C eps / tol = 5.5e-4)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_parity import BOX, VH, _helfand_case, make_universe  # noqa: E402

os.environ["TA_B200_HELFAND_FFT_THR"] = "-1"         # floor = -sum g^2: nothing is refined (read by ta_helfand_fft only)
eps = 2.0 ** -53
for kind, T, N in [("white", 3000, 40), ("smooth", 3000, 40), ("walk", 5000, 60), ("ramp", 2000, 3), ("white", 10000, 40),
                   ("smooth", 10000, 20), ("walk", 12000, 20), ("spike1", 3000, 40), ("offset_above", 3000, 24),
                   ("offset_below", 3000, 24), ("piecewise", 3000, 16), ("masses200", 5000, 64), ("white", 28999, 3),
                   ("smooth", 28999, 3)]:
    vel, pos, masses = _helfand_case(kind, T, N, seed=T + N)
    u = make_universe(pos, vel, masses=masses, dimensions=BOX)
    exact = VH(u.atoms, fft=False).run()            # the direct sums (K3): the comparator
    fast = VH(u.atoms, fft=True).run()              # S1 - 2 S2, un-refined
    assert exact.fft is False and fast.fft is True
    assert fast._ctx.helfand_fft_refined() == 0, fast._ctx.helfand_fft_refined()
    scale = 2 * exact.boltzmann * exact._vol_avg * exact.temp_avg * 3            # denom * D
    e, f = np.asarray(exact.results.visc_by_particle), np.asarray(fast.results.visc_by_particle)
    nk = (T - np.arange(T))[:, None]
    err = np.abs(f - e) * nk * scale                                              # un-normalised absolute error
    g = masses[None, :, None] * vel.astype(np.float64) * pos.astype(np.float64)
    tot = (g ** 2).sum(axis=(0, 2))[None, :]
    C = (err / (eps * tot))[1:]
    rel = (np.abs(f - e)[1:] / np.maximum(e[1:], 1e-300))
    print(f"{kind:7s} T={T:6d}: max C = {C.max():8.2f}   (worst relative error without refinement {rel.max():.2e})")
