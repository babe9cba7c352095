"""numpy restatement of the reference hot path (TEST INFRASTRUCTURE ONLY).

Each function states which reference lines it follows (paths relative to
/root/reference).  The arithmetic (operation order, dtype, normalisation)
is kept the same as the reference so that this module can serve as the
parity authority; see ``oracle/__init__.py`` for how it is pinned.

Array conventions are the reference's own: frame-major ``[T, N, D]`` float64
inputs, lag-major ``[T, N]`` per-particle outputs, ``[T]`` timeseries.
"""
from __future__ import annotations

import numpy as np

# MDAnalysis.units.constants["Boltzmann_constant"] (kJ mol^-1 K^-1).  The
# dependency is not vendored; SURVEY.md section 8(a8) confirms this value by
# reproducing docs/tutorials/helfand_dev_toy_system.ipynb:572 to all printed
# digits.  Used by transport_analysis/viscosity.py:139-142, :229-231.
BOLTZMANN_KJ_PER_MOL_K = 8.314462159e-3

_DIM_COLUMNS = {
    "x": [0],
    "y": [1],
    "z": [2],
    "xy": [0, 1],
    "xz": [0, 2],
    "yz": [1, 2],
    "xyz": [0, 1, 2],
}


def parse_dim_type(dim_str: str):
    """dim_type -> (column list, dimensionality).

    Follows transport_analysis/velocityautocorr.py:155-176 (duplicated at
    transport_analysis/viscosity.py:144-165), including the error text.
    The caller lower-cases first (velocityautocorr.py:133).
    """
    try:
        cols = _DIM_COLUMNS[dim_str]
    except KeyError:
        raise ValueError(
            "invalid dim_type: {} specified, please specify one of xyz, "
            "xy, xz, yz, x, y, z".format(dim_str)
        )
    return cols, len(cols)


# --------------------------------------------------------------------------
# tidynamics.acf  (third-party, pinned only as tidynamics>=1.0.0 in
# pyproject.toml:22 / setup.py:54; source not under /root/reference).
# Published algorithm of tidynamics 1.x `_correlation.py`, restated.
# Call sites: transport_analysis/velocityautocorr.py:211-213,
#             transport_analysis/tests/test_velocityautocorr.py:114.
# --------------------------------------------------------------------------
def _tidynamics_n_fft(n: int) -> int:
    """tidynamics' power-of-two selection: 2**ceil(log2(n+1)) for every n>=1."""
    e = int(np.ceil(np.log2(n + 1)))
    if n == 2 ** e:
        return n
    if n < 2 ** e:
        return 2 ** e
    return 2 ** (e + 1)


def _acf_1d(x: np.ndarray) -> np.ndarray:
    n = len(x)
    n_fft = _tidynamics_n_fft(n)
    padded = np.zeros(2 * n_fft)
    padded[:n] = x
    spec = np.fft.fft(padded)
    corr = np.fft.ifft(spec * spec.conj())[:n].real / (n - np.arange(n))
    return corr[:n]


def tidynamics_acf(data) -> np.ndarray:
    """Zero-padded FFT autocorrelation, summed over the last axis of [T, D]."""
    data = np.asarray(data)
    if data.ndim == 1:
        return _acf_1d(data)
    out = _acf_1d(data[:, 0])
    for j in range(1, data.shape[1]):
        out += _acf_1d(data[:, j])
    return out


def vacf_fft(velocities: np.ndarray):
    """FFT route.  Follows velocityautocorr.py:208-215 (`_conclude_fft`).

    velocities: [T, N, D] float64.  Returns (vacf_by_particle [T, N],
    timeseries [T]).
    """
    T, N, _ = velocities.shape
    by_particle = np.zeros((T, N))
    for n in range(N):
        by_particle[:, n] = tidynamics_acf(velocities[:, n, :])
    return by_particle, by_particle.mean(axis=1)


def vacf_windowed(velocities: np.ndarray):
    """Windowed route.  Follows velocityautocorr.py:217-238 (`_conclude_simple`):
    for every lag, product of the two shifted slabs, SUM over dims (:231),
    MEAN over the T-lag origins (:235), then mean over particles (:237).
    """
    T, N, _ = velocities.shape
    by_particle = np.zeros((T, N))
    for lag in range(T):
        prod = velocities[: T - lag, :, :] * velocities[lag:, :, :]
        by_particle[lag, :] = np.mean(np.sum(prod, axis=-1), axis=0)
    return by_particle, by_particle.mean(axis=1)


def helfand_msd(
    velocities: np.ndarray,
    positions: np.ndarray,
    masses: np.ndarray,
    volumes: np.ndarray,
    temp_avg: float = 300.0,
    boltzmann: float = BOLTZMANN_KJ_PER_MOL_K,
    lags=None,
):
    """Helfand-moment MSD.  Follows viscosity.py:201-233 (`_conclude`).

    Per lag >= 1: diff = m*v*x at the early origin minus m*v*x at the late one
    (evaluation order ((m*v)*x), :212-219), squared, MEAN over dims (:222),
    mean over origins (:226); row 0 stays 0; scale by 1/(2 kB <V> T) (:229-231);
    mean over particles (:233).

    ``lags`` (optional iterable) restricts the loop to a sample of lags -- used
    only by bench.py's bounded cpu_baseline leg; untouched rows stay 0.
    """
    T, N, _ = velocities.shape
    m = np.asarray(masses, dtype=np.float64).reshape((1, N, 1))
    vol_avg = np.average(volumes)
    by_particle = np.zeros((T, N))
    lag_iter = np.arange(1, T) if lags is None else lags
    for lag in lag_iter:
        diff = (
            m * velocities[:-lag, :, :] * positions[:-lag, :, :]
            - m * velocities[lag:, :, :] * positions[lag:, :, :]
        )
        by_particle[lag, :] = np.mean(np.square(diff).mean(axis=-1), axis=0)
    by_particle = by_particle / (2 * boltzmann * vol_avg * temp_avg)
    return by_particle, by_particle.mean(axis=1)


def polyfit_viscosity(timeseries: np.ndarray, fit_start: int, fit_end: int) -> float:
    """Slope of the viscosity function.  Follows viscosity.py:235-245: x runs
    over lagtimes = arange(1, T) while y is sliced from lag 0 -- kept as is."""
    lagtimes = np.arange(1, len(timeseries))
    return np.polyfit(
        lagtimes[fit_start:fit_end], timeseries[fit_start:fit_end], 1
    )[0]
