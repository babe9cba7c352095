"""Known answers the reference's own tests and notebooks hold for the hot path
(TEST INFRASTRUCTURE ONLY).

* `characteristic_poly`            -- closed-form VACF of the ramp v(t) = t,
  transport_analysis/tests/test_velocityautocorr.py:79-93.
* `characteristic_poly_helfand`    -- independent restatement of the Helfand
  loop for the ramp trajectory, transport_analysis/tests/test_viscosity.py:89-132.
* notebook vectors                 -- printed outputs of docs/tutorials.

The ramp sums are integers below 2**53, so an exact integer evaluation gives
the same float64 values as the reference's Python accumulation loop.
"""
from __future__ import annotations

import numpy as np

from .reference_numpy import BOLTZMANN_KJ_PER_MOL_K

# docs/tutorials/vacf_testing_examples.ipynb:90-99 -- VACF of the 10-frame
# ramp, dim_type="xyz" (last entry prints as rounding noise ~ -1e-14).
NOTEBOOK_VACF_T10_XYZ = np.array(
    [85.5, 80.0, 73.5, 66.0, 57.5, 48.0, 37.5, 26.0, 13.5, 0.0]
)

# docs/tutorials/helfand_dev_toy_system.ipynb:572,872 -- 10-frame ramp Helfand
# output of the notebook's older SUM-over-dims variant (notebook :243); the
# shipped module takes the MEAN over dims (viscosity.py:222), i.e. this / 3
# for dim_type="xyz".
NOTEBOOK_HELFAND_T10_SUMDIMS = np.array(
    [
        0.0,
        56426.98120794,
        192868.75919739,
        374484.83623557,
        583811.66540588,
        819319.38226769,
        1095007.68972109,
        1441040.89607656,
        1905422.10632966,
        2556706.56664059,
    ]
)


def frames_used(last: int, first: int = 0, step: int = 1) -> int:
    """Frame-count rule of test_velocityautocorr.py:80-82 (== len(range(first,
    last, step)) for the cases the reference tests)."""
    diff = last - first
    return int(diff // step + 1 if diff % step != 0 else diff / step)


def characteristic_poly(last: int, n_dim: int, first: int = 0, step: int = 1) -> np.ndarray:
    """C[k] = n_dim * sum_{x in range(first, last - k*step, step)} x*(x + k*step)
    / (frames_used - k), evaluated exactly in int64."""
    nf = frames_used(last, first, step)
    out = np.zeros(nf)
    for t in range(first, last, step):
        lag = t - first
        x = np.arange(first, last - lag, step, dtype=np.int64)
        idx = lag // step
        out[idx] = np.float64(int(np.sum(x * (x + lag)))) * n_dim / (nf - idx)
    return out


def ramp_trajectory(n_frames: int, first: int = 0, step: int = 1, stop=None):
    """Velocities/positions of the reference's step trajectory fixtures
    (test_velocityautocorr.py:46-72, test_viscosity.py:59-86): v = t and
    x = t^2/2 on every component, passed through float32 as an MDAnalysis
    Timestep would, for frames range(first, stop, step).
    Returns float64 arrays [T, 1, 3]."""
    stop = n_frames if stop is None else stop
    t = np.arange(n_frames, dtype=np.float64)[first:stop:step]
    v = t.astype(np.float32).astype(np.float64)
    x = (t * t / 2).astype(np.float32).astype(np.float64)
    vel = np.repeat(v[:, None, None], 3, axis=2)
    pos = np.repeat(x[:, None, None], 3, axis=2)
    return vel, pos


def characteristic_poly_helfand(
    velocities: np.ndarray,
    positions: np.ndarray,
    temp_avg: float = 300.0,
    mass: float = 16.0,
    vol_avg: float = 8.0,
    boltzmann: float = BOLTZMANN_KJ_PER_MOL_K,
) -> np.ndarray:
    """test_viscosity.py:89-132: mass * (v*x - v'*x') per lag (note the
    different association from the product code), squared, mean over dims,
    mean over origins, / (2 kB V T).  Inputs are the already dim-selected
    [T, 1, n_dim] arrays read back from the test universe (:118-120)."""
    nf = velocities.shape[0]
    result = np.zeros(nf)
    for lag in range(1, nf):
        diff = mass * (
            velocities[:-lag] * positions[:-lag] - velocities[lag:] * positions[lag:]
        )
        result[lag] = np.mean(np.square(diff).mean(axis=-1), axis=0)[0]
    return result / (2 * boltzmann * vol_avg * temp_avg)
