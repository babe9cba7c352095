"""The oracle against every golden vector / known answer the reference holds
for the hot path.  CPU only."""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_almost_equal, assert_array_equal

import oracle
from oracle.known_answers import ramp_trajectory

DIMS = [("xyz", 3), ("xy", 2), ("xz", 2), ("yz", 2), ("x", 1), ("y", 1), ("z", 1)]


def test_notebook_vacf_t10():
    vel, _ = ramp_trajectory(10)
    _, ts_w = oracle.vacf_windowed(vel)
    _, ts_f = oracle.vacf_fft(vel)
    assert_array_equal(ts_w, oracle.NOTEBOOK_VACF_T10_XYZ)          # windowed route is exact
    assert_allclose(ts_f, oracle.NOTEBOOK_VACF_T10_XYZ, atol=1e-12)  # notebook prints -1.07e-14 at the last lag


def test_notebook_helfand_t10():
    vel, pos = ramp_trajectory(10)
    _, ts = oracle.helfand_msd(vel, pos, np.array([16.0]), np.full(10, 8.0), 300.0)
    # notebook variant sums over dims; the shipped module takes the mean (viscosity.py:222)
    assert_allclose(ts * 3, oracle.NOTEBOOK_HELFAND_T10_SUMDIMS, rtol=1e-12)


@pytest.mark.parametrize("n_dim", [1, 2, 3])
def test_tidynamics_acf_vs_characteristic_poly(n_dim):
    # reference: tests/test_velocityautocorr.py:96-123 (decimal=4)
    vel, _ = ramp_trajectory(5001)
    got = oracle.tidynamics_acf(vel[:, 0, :n_dim])
    assert_almost_equal(got, oracle.characteristic_poly(5001, n_dim), decimal=4)


@pytest.mark.parametrize("dim,n_dim", DIMS)
def test_ramp_windowed_and_fft_full_and_sliced(dim, n_dim):
    # reference: TestAllDims, tests/test_velocityautocorr.py:331-360, :454-483
    cols, _ = oracle.parse_dim_type(dim)
    vel, _ = ramp_trajectory(5001, 10, 10, 1000)
    poly = oracle.characteristic_poly(1000, n_dim, first=10, step=10)
    assert_almost_equal(oracle.vacf_windowed(vel[:, :, cols])[1], poly, decimal=4)
    assert_almost_equal(oracle.vacf_fft(vel[:, :, cols])[1], poly, decimal=3)
    if dim in ("xyz", "x"):
        vel, _ = ramp_trajectory(1201)
        poly = oracle.characteristic_poly(1201, n_dim)
        assert_almost_equal(oracle.vacf_windowed(vel[:, :, cols])[1], poly, decimal=4)
        assert_almost_equal(oracle.vacf_fft(vel[:, :, cols])[1], poly, decimal=3)


@pytest.mark.parametrize("dim,n_dim", DIMS)
def test_ramp_helfand(dim, n_dim):
    # reference: tests/test_viscosity.py:167-208 (assert_allclose default rtol)
    cols, _ = oracle.parse_dim_type(dim)
    for args in ((801,), (5001, 10, 10, 1000)):
        vel, pos = ramp_trajectory(*args)
        _, ts = oracle.helfand_msd(vel[:, :, cols], pos[:, :, cols], np.array([16.0]),
                                   np.full(len(vel), 8.0), 300.0)
        expect = oracle.characteristic_poly_helfand(vel[:, :, cols], pos[:, :, cols])
        assert_allclose(ts, expect)
        assert ts[0] == 0.0


def test_dim_type_errors():
    for bad in ("foo", "bar", "yx", "zyx"):
        with pytest.raises(ValueError, match=f"invalid dim_type: {bad}"):
            oracle.parse_dim_type(bad)


def test_against_reference_run(golden):
    """Vectors produced by executing the reference's own modules
    (tests/golden/make_golden.py)."""
    vel, pos = golden["rand_vel"].astype(np.float64), golden["rand_pos"].astype(np.float64)
    vols = np.array([float(np.prod(b[:3].astype(np.float64))) for b in golden["rand_box"]])
    for dim, _ in DIMS:
        cols, _ = oracle.parse_dim_type(dim)
        bp, ts = oracle.vacf_windowed(vel[:, :, cols])
        assert_array_equal(ts, golden[f"rand_vacf_windowed_{dim}_ts"])
        assert_array_equal(bp, golden[f"rand_vacf_windowed_{dim}_bp"])
        bp, ts = oracle.vacf_fft(vel[:, :, cols])
        assert_array_equal(ts, golden[f"rand_vacf_fft_{dim}_ts"])
        bp, ts = oracle.helfand_msd(vel[:, :, cols], pos[:, :, cols], golden["rand_masses"], vols, 310.0)
        assert_allclose(ts, golden[f"rand_helfand_{dim}_ts"], rtol=1e-13)
        assert_allclose(bp, golden[f"rand_helfand_{dim}_bp"], rtol=1e-13)
    # sliced run on an atom subset
    sl = slice(3, 150, 4)
    _, ts = oracle.vacf_windowed(vel[sl, 1:5])
    assert_array_equal(ts, golden["rand_vacf_windowed_sliced_ts"])
    _, ts = oracle.helfand_msd(vel[sl, 1:5], pos[sl, 1:5], golden["rand_masses"][1:5], vols[sl], 300.0)
    assert_allclose(ts, golden["rand_helfand_sliced_ts"], rtol=1e-13)
    assert_allclose(oracle.polyfit_viscosity(ts, 5, 30), golden["rand_helfand_sliced_viscosity"], rtol=1e-10)


def test_fft_route_equals_windowed_route():
    # reference: TestVACFFFT, tests/test_velocityautocorr.py:297-315 (decimal=4)
    rng = np.random.default_rng(3)
    vel = rng.standard_normal((300, 4, 3))
    bw, tw = oracle.vacf_windowed(vel)
    bf, tf = oracle.vacf_fft(vel)
    assert_almost_equal(tw, tf, decimal=10)
    assert_almost_equal(bw, bf, decimal=10)
