"""Parity of the CUDA path against the oracle, through the public classes and
the C ABI (ctypes).  FP64 bar: |gpu - ref| <= 1e-10 (|ref| + max|ref|) for the
VACF routes (SURVEY.md section 7 tolerance definition; BASELINE.json states
rel. 1e-10), plain rtol 1e-10 for Helfand; FP32 mode: 1e-5 on the same norm.
"""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_almost_equal, assert_approx_equal

import oracle
from oracle.known_answers import ramp_trajectory
from _util import assert_close_normwise
from transport_analysis_b200 import _lib
from transport_analysis_b200.synthetic import make_universe, random_trajectory
from transport_analysis_b200.velocityautocorr import VelocityAutocorr as VACF
from transport_analysis_b200.viscosity import ViscosityHelfand as VH

pytestmark = pytest.mark.gpu

DIMS = [("xyz", 3), ("xy", 2), ("xz", 2), ("yz", 2), ("x", 1), ("y", 1), ("z", 1)]
TOL64 = 1e-10
TOL32 = 1e-5
BOX = [20.0, 21.0, 19.0, 90.0, 90.0, 90.0]


def _f64(a):
    return np.asarray(a, dtype=np.float64)


@pytest.fixture(scope="module")
def rand_u():
    vel, pos = random_trajectory(700, 37, seed=11, with_positions=True, rho=0.9)
    masses = np.random.default_rng(0).choice([1.008, 12.011, 15.999], 37)
    return make_universe(pos, vel, masses=masses, dimensions=BOX), vel, pos, masses


@pytest.fixture(scope="module")
def step_u():
    # the reference's step trajectory: v = t, x = t^2/2, mass 16, volume 8
    t = np.arange(5001, dtype=np.float64)
    v = np.repeat(t[:, None, None], 3, axis=2)
    x = np.repeat((t * t / 2)[:, None, None], 3, axis=2)
    return make_universe(x, v, masses=[16.0], dimensions=[2, 2, 2, 90, 90, 90])


# ------------------------------------------------------------------ VACF
@pytest.mark.parametrize("dim,n_dim", DIMS)
@pytest.mark.parametrize("fft", [True, False])
def test_vacf_random_all_dims(rand_u, dim, n_dim, fft):
    u, vel, _, _ = rand_u
    cols, _ = oracle.parse_dim_type(dim)
    ref_bp, ref_ts = (oracle.vacf_fft if fft else oracle.vacf_windowed)(_f64(vel)[:, :, cols])
    v = VACF(u.atoms, dim_type=dim, fft=fft).run()
    assert v.results.timeseries.dtype == np.float64
    assert v.results.vacf_by_particle.shape == (v.n_frames, v.n_particles)
    assert_close_normwise(v.results.timeseries, ref_ts, TOL64, "timeseries")
    assert_close_normwise(v.results.vacf_by_particle, ref_bp, TOL64, "by particle")


@pytest.mark.parametrize("fft", [True, False])
def test_vacf_sliced_subset_per_frame_path(rand_u, fft):
    u, vel, _, _ = rand_u
    ag = u.atoms[[3, 5, 6, 20, 36]]          # non-contiguous -> per-frame pinned-slab path
    sl = slice(5, 690, 3)
    ref_bp, ref_ts = (oracle.vacf_fft if fft else oracle.vacf_windowed)(_f64(vel)[sl][:, [3, 5, 6, 20, 36]])
    v = VACF(ag, fft=fft).run(start=5, stop=690, step=3)
    assert not v._stager.bulk_done
    assert_close_normwise(v.results.timeseries, ref_ts, TOL64)
    assert_close_normwise(v.results.vacf_by_particle, ref_bp, TOL64)
    assert_allclose(v.times, np.arange(700.0)[sl])
    # contiguous subset -> bulk path, same numbers as the per-frame path
    ag2 = u.atoms[3:21]
    vb = VACF(ag2, fft=fft).run(start=5, stop=690, step=3)
    assert vb._stager.bulk_done
    ref_bp, ref_ts = (oracle.vacf_fft if fft else oracle.vacf_windowed)(_f64(vel)[sl][:, 3:21])
    assert_close_normwise(vb.results.timeseries, ref_ts, TOL64)


@pytest.mark.parametrize("dim,n_dim", DIMS)
def test_vacf_step_trajectory_known_answer(step_u, dim, n_dim):
    # reference: TestAllDims (tests/test_velocityautocorr.py:331-360, :454-483)
    poly = oracle.characteristic_poly(5001, n_dim)
    v_fft = VACF(step_u.atoms, dim_type=dim, fft=True).run()
    assert_almost_equal(v_fft.results.timeseries, poly, decimal=3)
    if dim in ("xyz", "y"):
        v_simple = VACF(step_u.atoms, dim_type=dim, fft=False).run()
        assert_almost_equal(v_simple.results.timeseries, poly, decimal=4)
        sd_expected = __import__("scipy").integrate.simpson(y=poly, x=range(5001)) / n_dim
        assert_approx_equal(v_simple.self_diffusivity_gk(), sd_expected, significant=8)
        assert_approx_equal(v_fft.self_diffusivity_gk(), sd_expected, significant=8)
    poly = oracle.characteristic_poly(1000, n_dim, first=10, step=10)
    for fft, dec in ((True, 3), (False, 4)):
        v = VACF(step_u.atoms, dim_type=dim, fft=fft).run(start=10, stop=1000, step=10)
        assert_almost_equal(v.results.timeseries, poly, decimal=dec)


def test_notebook_vector_t10():
    vel, _ = ramp_trajectory(10)
    u = make_universe(None, vel)
    assert_allclose(VACF(u.atoms, fft=True).run().results.timeseries, oracle.NOTEBOOK_VACF_T10_XYZ, atol=1e-11)
    assert_allclose(VACF(u.atoms, fft=False).run().results.timeseries, oracle.NOTEBOOK_VACF_T10_XYZ, atol=1e-13)


def test_against_reference_run(golden):
    vel, pos = golden["rand_vel"], golden["rand_pos"]
    u = make_universe(pos, vel, masses=golden["rand_masses"], dimensions=golden["rand_box"])
    for dim, _ in DIMS:
        v = VACF(u.atoms, dim_type=dim, fft=False).run()
        assert_close_normwise(v.results.timeseries, golden[f"rand_vacf_windowed_{dim}_ts"], TOL64)
        assert_close_normwise(v.results.vacf_by_particle, golden[f"rand_vacf_windowed_{dim}_bp"], TOL64)
        v = VACF(u.atoms, dim_type=dim, fft=True).run()
        assert_close_normwise(v.results.timeseries, golden[f"rand_vacf_fft_{dim}_ts"], TOL64)
        assert_close_normwise(v.results.vacf_by_particle, golden[f"rand_vacf_fft_{dim}_bp"], TOL64)
        h = VH(u.atoms, temp_avg=310.0, dim_type=dim).run()
        assert_allclose(h.results.timeseries, golden[f"rand_helfand_{dim}_ts"], rtol=TOL64)
        assert_allclose(h.results.visc_by_particle, golden[f"rand_helfand_{dim}_bp"], rtol=TOL64)
    h = VH(u.atoms[1:5], linear_fit_window=(5, 30)).run(start=3, stop=150, step=4)
    assert_allclose(h.results.timeseries, golden["rand_helfand_sliced_ts"], rtol=TOL64)
    assert_allclose(h.results.viscosity, golden["rand_helfand_sliced_viscosity"], rtol=1e-9)
    v = VACF(u.atoms, fft=False).run()
    assert_allclose(v.self_diffusivity_gk(), golden["rand_gk"], rtol=1e-10)
    assert_allclose(v.self_diffusivity_gk_odd(start=0, stop=159), golden["rand_gk_odd"], rtol=1e-10)
    assert_allclose(v.self_diffusivity_gk(start=2, stop=100, step=3), golden["rand_gk_sliced"], rtol=1e-10)


@pytest.mark.parametrize("T,N", [(1, 1), (2, 3), (3, 1), (15, 2), (16, 2), (17, 5), (33, 1), (257, 3), (1000, 300)])
def test_edge_sizes(T, N):
    vel, pos = random_trajectory(T, N, seed=T * 31 + N, with_positions=True)
    u = make_universe(pos, vel, masses=np.linspace(1, 2, N), dimensions=BOX)
    rw_bp, rw_ts = oracle.vacf_windowed(_f64(vel))
    rf_bp, rf_ts = oracle.vacf_fft(_f64(vel))
    assert_close_normwise(VACF(u.atoms, fft=False).run().results.vacf_by_particle, rw_bp, TOL64)
    assert_close_normwise(VACF(u.atoms, fft=True).run().results.vacf_by_particle, rf_bp, TOL64)
    vols = np.full(T, float(np.prod(np.float32(BOX[:3]).astype(np.float64))))
    rh_bp, rh_ts = oracle.helfand_msd(_f64(vel), _f64(pos), np.linspace(1, 2, N), vols, 300.0)
    h = VH(u.atoms).run()
    assert_allclose(h.results.visc_by_particle, rh_bp, rtol=TOL64)
    assert_allclose(h.results.timeseries, rh_ts, rtol=TOL64)


# ------------------------------------------------------------------ Helfand
HELFAND_ROUTES = ["auto", False]      # the default (FFT + exact refinement) and the direct lag sums


@pytest.mark.parametrize("route", HELFAND_ROUTES)
@pytest.mark.parametrize("dim,n_dim", DIMS)
def test_helfand_random_all_dims(rand_u, dim, n_dim, route):
    u, vel, pos, masses = rand_u
    cols, _ = oracle.parse_dim_type(dim)
    vols = np.full(700, float(np.prod(np.float32(BOX[:3]).astype(np.float64))))
    ref_bp, ref_ts = oracle.helfand_msd(_f64(vel)[:, :, cols], _f64(pos)[:, :, cols], masses, vols, 300.0)
    h = VH(u.atoms, dim_type=dim, fft=route).run()
    assert h.results.timeseries[0] == 0.0 and np.all(h.results.visc_by_particle[0] == 0.0)
    assert_allclose(h.results.timeseries, ref_ts, rtol=TOL64)
    assert_allclose(h.results.visc_by_particle, ref_bp, rtol=TOL64)


@pytest.mark.parametrize("route", HELFAND_ROUTES)
@pytest.mark.parametrize("dim,n_dim", [("xyz", 3), ("yz", 2), ("z", 1)])
def test_helfand_step_trajectory_known_answer(step_u, dim, n_dim, route):
    # reference: tests/test_viscosity.py:167-208 (assert_allclose, default rtol 1e-7); here at the FP64 bar
    cols, _ = oracle.parse_dim_type(dim)
    vel, pos = ramp_trajectory(5001, 10, 10, 1000)
    expect = oracle.characteristic_poly_helfand(vel[:, :, cols], pos[:, :, cols])
    h = VH(step_u.atoms, dim_type=dim, fft=route).run(start=10, stop=1000, step=10)
    assert_allclose(h.results.timeseries, expect, rtol=TOL64)
    if dim == "xyz":
        vel, pos = ramp_trajectory(5001)
        expect = oracle.characteristic_poly_helfand(vel, pos)
        assert_allclose(VH(step_u.atoms, fft=route).run().results.timeseries, expect, rtol=TOL64)


@pytest.mark.parametrize("route", HELFAND_ROUTES)
def test_notebook_helfand_t10(route):
    vel, pos = ramp_trajectory(10)
    u = make_universe(pos, vel, masses=[16.0], dimensions=[2, 2, 2, 90, 90, 90])
    assert_allclose(VH(u.atoms, fft=route).run().results.timeseries * 3, oracle.NOTEBOOK_HELFAND_T10_SUMDIMS, rtol=1e-10)


# ------------------------------------------------------------------ device-side Green-Kubo / fit (kernel K7)
def test_green_kubo_and_linear_fit_on_the_device(rand_u):
    """postprocess='device': trapezoid integral, running integral and the least-squares slope are computed on the GPU from
    the timeseries that is still there; against the reference's scipy / numpy calls on the host timeseries."""
    from scipy import integrate

    u, vel, pos, masses = rand_u
    host = VACF(u.atoms, fft=True).run()
    dev = VACF(u.atoms, fft=True, postprocess="device").run()
    assert np.array_equal(host.results.timeseries, dev.results.timeseries)
    for kw in ({}, {"start": 2, "stop": 100, "step": 3}, {"start": 5, "stop": 0, "step": 1}, {"start": 0, "stop": 699, "step": 2}):
        assert_allclose(dev.self_diffusivity_gk(**kw), host.self_diffusivity_gk(**kw), rtol=1e-12)
        t_d, r_d = dev.running_integral(**kw)
        t_h, r_h = host.running_integral(**kw)
        assert_allclose(t_d, t_h)
        assert_allclose(r_d, r_h, rtol=1e-12, atol=1e-13 * np.abs(r_h).max())
    assert_allclose(dev.self_diffusivity_gk_odd(stop=699), host.self_diffusivity_gk_odd(stop=699), rtol=1e-13)   # host both
    # a timeseries longer than one scan tile of K7 (1,024 points)
    v2, _ = random_trajectory(2500, 3, seed=4, rho=0.95)
    d2 = VACF(make_universe(None, v2).atoms, postprocess="device").run()
    want = integrate.trapezoid(d2.results.timeseries, d2.times) / 3
    assert_allclose(d2.self_diffusivity_gk(), want, rtol=1e-12)
    assert_allclose(d2.running_integral()[1], integrate.cumulative_trapezoid(d2.results.timeseries, d2.times, initial=0) / 3,
                    rtol=1e-12, atol=1e-13)
    for window in ((5, 30), (20, 650), (0, 699)):
        h_host = VH(u.atoms, linear_fit_window=window).run()
        h_dev = VH(u.atoms, linear_fit_window=window, postprocess="device").run()
        assert_allclose(h_dev.results.viscosity, h_host.results.viscosity, rtol=1e-10)
    with pytest.raises(ValueError, match="postprocess"):
        VACF(u.atoms, postprocess="nowhere")


# ------------------------------------------------------------------ C ABI direct
def test_c_abi_f64_source_lag_major_and_errors():
    rng = np.random.default_rng(21)
    T, N = 300, 9
    vel = rng.standard_normal((T, N, 3))
    ctx = _lib.Context([0])
    with pytest.raises(_lib.BackendError, match="ta_stage_begin has not been called"):
        ctx.T = T
        ctx.vacf_fft()
    ctx.stage_begin(T, N, [0, 2], np.float64, 1, None, "fp64")
    with pytest.raises(_lib.BackendError, match="frames were staged"):
        ctx.vacf_fft()
    ctx.stage_bulk([vel])
    ctx.stage_end()
    ts = ctx.vacf_fft()
    ref_bp, ref_ts = oracle.vacf_fft(vel[:, :, [0, 2]])
    assert_close_normwise(ts, ref_ts, TOL64)
    assert_close_normwise(ctx.fetch_by_particle(), ref_bp, TOL64)
    lm = ctx.fetch_by_particle(lag_major_copy=True)
    assert lm.flags.c_contiguous
    assert np.array_equal(lm, ctx.fetch_by_particle())
    assert np.array_equal(ctx.fetch_by_particle(2, 4), ctx.fetch_by_particle()[:, 2:6])
    ts_w = ctx.vacf_windowed()            # same staged data, other route
    assert_close_normwise(ts_w, oracle.vacf_windowed(vel[:, :, [0, 2]])[1], TOL64)
    with pytest.raises(_lib.BackendError, match="n_fields == 2"):
        ctx.helfand(np.ones(T), 1.0, 300.0)
    assert ctx.launch_count() > 0
    info = ctx.fft_plan_info()
    assert info["H"] >= T // 2 and int(np.prod(info["radices"])) == info["H"]
    ctx.close()


def test_results_are_bit_reproducible(rand_u):
    u = rand_u[0]
    a = VACF(u.atoms).run().results.vacf_by_particle
    b = VACF(u.atoms).run().results.vacf_by_particle
    assert np.array_equal(a, b)


def test_fp32_mode(rand_u, monkeypatch):
    u, vel, pos, masses = rand_u
    for fft in (True, False):
        ref_bp, ref_ts = (oracle.vacf_fft if fft else oracle.vacf_windowed)(_f64(vel))
        v = VACF(u.atoms, fft=fft, precision="fp32").run()
        assert_close_normwise(v.results.timeseries, ref_ts, TOL32)
        assert_close_normwise(v.results.vacf_by_particle, ref_bp, TOL32)
    # one context, FP64 then FP32 then FP64: the series buffers are rebuilt in the arithmetic type
    a = VACF(u.atoms, fft=True)
    first = np.array(a.run().results.timeseries)
    a.precision = "fp32"
    assert_close_normwise(a.run().results.timeseries, first, TOL32)
    a.precision = "fp64"
    assert np.array_equal(a.run().results.timeseries, first)


def test_lazy_by_particle(rand_u):
    u = rand_u[0]
    eager = VACF(u.atoms).run().results.vacf_by_particle
    lazy = VACF(u.atoms, max_eager_bytes=0).run().results.vacf_by_particle
    assert lazy.shape == eager.shape
    assert np.array_equal(np.asarray(lazy), eager)
    assert np.array_equal(lazy.particles(4, 9), eager[:, 4:9])


def test_lazy_handle_of_an_earlier_run_goes_stale(rand_u):
    """A per-particle result left on the GPU belongs to the device buffers of its run: once run() is called again an
    unread handle refuses to return values (they would be the new run's), a handle that was read keeps its copy."""
    u, _, _, _ = rand_u
    a = VACF(u.atoms, fft=True, max_eager_bytes=0)
    read = a.run().results.vacf_by_particle
    kept = np.array(read)
    unread = a.run(start=5, stop=400).results.vacf_by_particle
    assert np.array_equal(np.asarray(read), kept) and read.shape == kept.shape
    third = a.run().results.vacf_by_particle
    with pytest.raises(RuntimeError, match="later run"):
        np.asarray(unread)
    assert np.array_equal(np.asarray(third), kept)
    assert np.array_equal(third.mean(axis=1), kept.mean(axis=1))          # ndarray methods work on the handle


# ------------------------------------------------------------------ K1 fast path (H = 256 R1, k1_fast.cuh)
@pytest.mark.parametrize("T,N,dim", [(1600, 5, "xyz"), (2000, 70, "xyz"), (2047, 3, "x"), (3000, 9, "yz"), (4000, 4, "xyz"),
                                     (5000, 333, "xyz"), (5001, 2, "xy"), (6000, 3, "z"), (8192, 2, "xyz"),
                                     (10000, 301, "xyz"), (10240, 1, "xyz"), (10241, 3, "xz"), (12000, 150, "xyz"),
                                     (12288, 2, "y")])
def test_fft_fast_path_vs_oracle_and_general_kernel(T, N, dim, monkeypatch):
    """Every R1 instantiation of the three-pass kernel (4 ... 24): against the oracle, and against the general
    mixed-radix kernel (TA_B200_K1_PATH=general, read when a context is created) on the same data.
    N > 148 makes CTAs take several particles (prefetch-buffer hand-over across particles)."""
    vel, _ = random_trajectory(T, N, seed=T + N, rho=0.8)
    u = make_universe(None, vel)
    cols, _ = oracle.parse_dim_type(dim)
    v = VACF(u.atoms, dim_type=dim, fft=True).run()
    plan = v._ctx.fft_plan_info()
    assert plan["radices"][1:] == [16, 16] and plan["H"] == 256 * plan["radices"][0], plan
    n_or = min(N, 6)                                   # the oracle is a Python loop over particles
    ref_bp, _ = oracle.vacf_fft(_f64(vel)[:, :n_or, cols])
    assert_close_normwise(v.results.vacf_by_particle[:, :n_or], ref_bp, TOL64, "fast vs oracle, by particle")
    assert_close_normwise(v.results.timeseries, v.results.vacf_by_particle.mean(axis=1), 1e-13, "timeseries = particle mean")
    again = VACF(u.atoms, dim_type=dim, fft=True).run()
    assert np.array_equal(again.results.timeseries, v.results.timeseries)          # bit-reproducible
    assert np.array_equal(again.results.vacf_by_particle, v.results.vacf_by_particle)
    monkeypatch.setenv("TA_B200_K1_PATH", "general")
    g = VACF(u.atoms, dim_type=dim, fft=True).run()
    assert g._ctx.fft_plan_info()["radices"][1:] != [16, 16]
    assert_close_normwise(g.results.vacf_by_particle, v.results.vacf_by_particle, 1e-11, "fast vs general kernel")
    assert_close_normwise(g.results.timeseries, v.results.timeseries, 1e-11, "fast vs general kernel, timeseries")


@pytest.mark.parametrize("T,N,dim", [(10000, 700, "xyz"), (8000, 450, "xz"), (6200, 333, "y")])
def test_tensor_memory_output_stage_vs_global_one_and_across_launches(T, N, dim, monkeypatch):
    """R1 = 16 / 20 in FP64: the output stage keeps the parked V_0, the particle sums and the 1/(L(T-k)) table in tensor
    memory (tcgen05.st / ld on each thread's own columns).  Against the build that sends them through L2
    (TA_B200_K1_PATH=notmem): per-particle rows bit for bit; the timeseries too when one launch does the shard, and up to
    the order of the sums when the compute call launches once per staging chunk (each launch ADDS its sums to the global
    partial rows)."""
    vel, _ = random_trajectory(T, N, seed=T + N, rho=0.7)
    u = make_universe(None, vel)
    v = VACF(u.atoms, dim_type=dim, fft=True).run()
    assert v._ctx.fft_plan_info()["radices"][0] in (16, 20) and v._ctx.fft_plan_info()["tmem"]
    cols, _ = oracle.parse_dim_type(dim)
    ref_bp, _ = oracle.vacf_fft(_f64(vel)[:, :4, cols])
    assert_close_normwise(v.results.vacf_by_particle[:, :4], ref_bp, TOL64, "tensor-memory build vs oracle")
    assert_close_normwise(v.results.timeseries, v.results.vacf_by_particle.mean(axis=1), 1e-13, "timeseries = particle mean")
    monkeypatch.setenv("TA_B200_BULK_CHUNK", "148")
    c = VACF(u.atoms, dim_type=dim, fft=True).run()
    assert c._ctx.launch_count() >= 2 * (-(-N // 148))
    assert np.array_equal(c.results.vacf_by_particle, v.results.vacf_by_particle)
    assert_allclose(c.results.timeseries, v.results.timeseries, rtol=1e-13, atol=1e-15)
    assert_close_normwise(c.results.timeseries, c.results.vacf_by_particle.mean(axis=1), 1e-13, "chunked: timeseries = particle mean")
    monkeypatch.delenv("TA_B200_BULK_CHUNK")
    monkeypatch.setenv("TA_B200_K1_PATH", "notmem")
    g = VACF(u.atoms, dim_type=dim, fft=True).run()
    assert not g._ctx.fft_plan_info()["tmem"]
    assert np.array_equal(g.results.vacf_by_particle, v.results.vacf_by_particle)
    assert np.array_equal(g.results.timeseries, v.results.timeseries)


@pytest.mark.parametrize("T,want", [(1000, None), (2000, [4, 16, 16]), (5000, [10, 16, 16]), (10000, [20, 16, 16]),
                                    (12000, [24, 16, 16]), (13000, None)])
def test_fft_default_kernel_choice(T, want):
    """Three-pass radix-16 kernel where it has an instantiation (T <= 12,288), else the general kernel; both precisions."""
    vel, _ = random_trajectory(T, 2, seed=T, rho=0.5)
    ref_bp, ref_ts = oracle.vacf_fft(_f64(vel))
    for precision, tol in (("fp64", TOL64), ("fp32", TOL32)):
        v = VACF(make_universe(None, vel).atoms, fft=True, precision=precision).run()
        rad = v._ctx.fft_plan_info()["radices"]
        if want is None:
            assert not (len(rad) == 3 and rad[1:] == [16, 16]), rad
        else:
            assert rad == want, rad
        cut = None if precision == "fp64" else -8          # FP32: lags with at least 8 origins (DESIGN.md "FP32 mode")
        assert_close_normwise(v.results.vacf_by_particle[:cut], ref_bp[:cut], tol, f"T={T} {precision}")
        assert_close_normwise(v.results.timeseries[:cut], ref_ts[:cut], tol, f"T={T} {precision}")


@pytest.mark.parametrize("T,N,dim", [(19000, 3, "xyz"), (20001, 150, "xyz"), (30000, 2, "y"), (65536, 2, "xz")])
def test_fft_route_beyond_shared_memory(T, N, dim):
    """Trajectories whose FFT buffer does not fit in shared memory (T > ~19,000 in FP64) run the general kernel on a
    per-CTA global work area: the FFT route never refuses a length.  Against the oracle; N > 148 reuses work areas."""
    vel, _ = random_trajectory(T, N, seed=T % 97, rho=0.9)
    u = make_universe(None, vel)
    cols, _ = oracle.parse_dim_type(dim)
    v = VACF(u.atoms, dim_type=dim, fft=True).run()
    info = v._ctx.fft_plan_info()
    assert info["H"] >= (T + 1) // 2 and info["smem_bytes"] < 100_000, info
    n_or = min(N, 3)
    ref_bp, _ = oracle.vacf_fft(_f64(vel)[:, :n_or, cols])
    assert_close_normwise(v.results.vacf_by_particle[:, :n_or], ref_bp, TOL64, f"T={T} by particle")
    assert_close_normwise(v.results.timeseries, v.results.vacf_by_particle.mean(axis=1), 1e-13, "timeseries = particle mean")
    if N <= 3:
        _, ref_ts = oracle.vacf_fft(_f64(vel)[:, :, cols])
        assert_close_normwise(v.results.timeseries, ref_ts, TOL64, f"T={T} timeseries")


def test_fft_fast_path_ramp_known_answer():
    """The reference's step trajectory (v = t, 5001 frames) takes the fast path (R1 = 10)."""
    t = np.arange(5001, dtype=np.float64)
    v = np.repeat(t[:, None, None], 3, axis=2)
    u = make_universe(None, v)
    a = VACF(u.atoms, fft=True).run()
    assert a._ctx.fft_plan_info()["radices"] == [10, 16, 16]
    poly = oracle.characteristic_poly(5001, 3)
    assert_almost_equal(a.results.timeseries, poly, decimal=3)     # the reference's own bar (tests :454-469)
    assert_close_normwise(a.results.timeseries, poly, TOL64)


# ------------------------------------------------------------------ size-independent properties at BASELINE sizes
def test_properties_at_config_sizes():
    """Config 1 size (1,000 x 5,000): the oracle is too slow for the full set,
    so check (a) 16 sampled particles against the oracle, (b) lag 0 equals the
    mean square, (c) the timeseries is the mean of the per-particle array,
    (d) FFT route == windowed route, (e) acf(a x) = a^2 acf(x)."""
    T, N = 5000, 1000
    vel, _ = random_trajectory(T, N, seed=5)
    u = make_universe(None, vel)
    v = VACF(u.atoms, fft=True).run()
    bp, ts = v.results.vacf_by_particle, v.results.timeseries
    pick = np.random.default_rng(0).choice(N, 16, replace=False)
    ref_bp, _ = oracle.vacf_fft(_f64(vel)[:, pick])
    assert_close_normwise(bp[:, pick], ref_bp, TOL64)
    assert_allclose(bp[0], (_f64(vel) ** 2).sum(axis=2).mean(axis=0), rtol=1e-12)
    assert_allclose(ts, bp.mean(axis=1), rtol=1e-12, atol=1e-13 * np.abs(ts).max())
    w = VACF(u.atoms[:64], fft=False).run()
    assert_close_normwise(w.results.vacf_by_particle, bp[:, :64], TOL64)
    u2 = make_universe(None, vel * np.float32(2.0))
    v2 = VACF(u2.atoms[:256], fft=True).run()
    assert_allclose(v2.results.vacf_by_particle, 4.0 * bp[:, :256], rtol=1e-12, atol=1e-12 * np.abs(bp).max())


def test_properties_at_config4_full_size():
    """BASELINE config 4 at full size (100,000 particles x 10,000 frames, the bench workload): the oracle cannot run it, so
    (a) sampled particles against the oracle, (b) lag 0 of sampled particles equals their mean square, (c) the timeseries
    equals the mean over ALL per-particle rows (fetched range by range from the device), (d) a scaled sub-range gives the
    scaled result, (e) a second run returns the same bits.  Needs ~25 GB of host memory and one GPU with 40 GB free."""
    import bench

    T, N = 10000, 100000
    try:
        vel = np.empty((T, N, 3), dtype=np.float32)
    except MemoryError:
        pytest.skip("not enough host memory for the full-size trajectory")
    bench.fill_random_f32(vel, seed=99)
    u = make_universe(None, vel)
    v = VACF(u.atoms, fft=True).run()
    assert v._ctx.fft_plan_info()["radices"] == [20, 16, 16]
    ts, lazy = v.results.timeseries, v.results.vacf_by_particle       # 8 GB: stays on the device behind a lazy handle
    assert lazy.shape == (T, N)
    pick = [0, 1, 31337, 49999, 50000, 77777, 99998, 99999]
    ref_bp, _ = oracle.vacf_fft(_f64(vel[:, pick]))
    got = np.stack([lazy.particles(p, p + 1)[:, 0] for p in pick], axis=1)
    assert_close_normwise(got, ref_bp, TOL64, "sampled particles vs oracle")
    assert_allclose(got[0], (_f64(vel[:, pick]) ** 2).sum(axis=2).mean(axis=0), rtol=1e-12)
    total = np.zeros(T)
    for a in range(0, N, 5000):
        total += lazy.particles(a, a + 5000).sum(axis=1)
    assert_allclose(ts, total / N, rtol=1e-11, atol=1e-13 * np.abs(ts).max())
    again = VACF(u.atoms, fft=True).run()
    assert np.array_equal(again.results.timeseries, ts)
    sub = vel[:, 1000:1512] * np.float32(2.0)
    v2 = VACF(make_universe(None, sub).atoms, fft=True).run()
    assert_allclose(v2.results.vacf_by_particle, 4.0 * lazy.particles(1000, 1512), rtol=1e-12, atol=1e-12 * np.abs(ref_bp).max())


def test_properties_at_config2_and_config3_full_size():
    """BASELINE config 2 (windowed VACF, 1,000 x 2,000) and config 3 (Helfand, 10,000 x 5,000) at full size: sampled
    particles against the oracle, windowed == FFT route on all particles, Helfand row 0 == 0, timeseries == particle mean,
    g -> 2 g gives 4 x the Helfand MSD, opt-in route (FFT + exact refinement) == default direct lag sums to 1e-10 everywhere."""
    import bench

    # ---- config 2
    T, N = 2000, 1000
    vel = np.empty((T, N, 3), dtype=np.float32)
    bench.fill_random_f32(vel, seed=7)
    u = make_universe(None, vel)
    w = VACF(u.atoms, fft=False).run()
    f = VACF(u.atoms, fft=True).run()
    assert_close_normwise(w.results.vacf_by_particle, f.results.vacf_by_particle, TOL64, "windowed vs FFT, all particles")
    pick = [0, 499, 500, 999]
    ref_bp, _ = oracle.vacf_windowed(_f64(vel[:, pick]))
    assert_close_normwise(w.results.vacf_by_particle[:, pick], ref_bp, TOL64, "windowed vs oracle")
    assert_allclose(w.results.timeseries, w.results.vacf_by_particle.mean(axis=1), rtol=1e-12,
                    atol=1e-13 * np.abs(w.results.timeseries).max())
    # ---- config 3
    T, N = 5000, 10000
    vel = np.empty((T, N, 3), dtype=np.float32)
    pos = np.empty((T, N, 3), dtype=np.float32)
    bench.fill_random_f32(vel, seed=8)
    bench.fill_random_f32(pos, seed=9)
    pos *= np.float32(10.0)
    masses = np.random.default_rng(3).choice([1.008, 12.011, 15.999], N)
    u = make_universe(pos, vel, masses=masses, dimensions=BOX)
    h = VH(u.atoms, fft=True).run()
    bp, ts = np.asarray(h.results.visc_by_particle), h.results.timeseries
    assert ts[0] == 0.0 and np.all(bp[0] == 0.0)
    assert_allclose(ts, bp.mean(axis=1), rtol=1e-12)
    pick = [0, 1, 5000, 9999]
    vols = np.full(T, float(np.prod(np.float32(BOX[:3]).astype(np.float64))))
    lags = [1, 2, 1250, 2500, 3750, 4998, 4999]
    ref_bp, _ = oracle.helfand_msd(_f64(vel[:, pick]), _f64(pos[:, pick]), masses[pick], vols, 300.0, lags=lags)
    assert_allclose(bp[lags][:, pick], ref_bp[lags], rtol=TOL64)
    hd = VH(u.atoms, fft=False).run()
    assert h.fft is True and hd.fft is False
    assert_allclose(hd.results.timeseries, ts, rtol=TOL64)
    assert_allclose(np.asarray(hd.results.visc_by_particle), bp, rtol=TOL64)
    h2 = VH(make_universe(pos[:, :256], vel[:, :256] * np.float32(2.0), masses=masses[:256], dimensions=BOX).atoms, fft=True).run()
    assert_allclose(np.asarray(h2.results.visc_by_particle), 4.0 * bp[:, :256], rtol=1e-12)


def test_multi_gpu_sharding_matches_single():
    n = _lib.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    vel, pos = random_trajectory(400, 101, seed=9, with_positions=True)
    u = make_universe(pos, vel, masses=np.ones(101), dimensions=BOX)
    one = VACF(u.atoms, devices=[0]).run()
    two = VACF(u.atoms, devices=list(range(min(n, 4)))).run()
    assert np.array_equal(one.results.vacf_by_particle, two.results.vacf_by_particle)
    assert_allclose(one.results.timeseries, two.results.timeseries, rtol=1e-13, atol=1e-14)
    ref_bp, ref_ts = oracle.vacf_fft(_f64(vel))
    assert_close_normwise(two.results.timeseries, ref_ts, TOL64, "NCCL-reduced timeseries vs oracle")
    assert_close_normwise(two.results.vacf_by_particle, ref_bp, TOL64)
    _, ref_h = oracle.helfand_msd(_f64(vel), _f64(pos), np.ones(101), np.full(400, np.prod(BOX[:3])), 300.0)
    for route in (False, True):
        h1 = VH(u.atoms, devices=[0], fft=route).run()
        h2 = VH(u.atoms, devices=[0, 1], fft=route).run()
        assert np.array_equal(h1.results.visc_by_particle, h2.results.visc_by_particle)
        assert_allclose(h1.results.timeseries, h2.results.timeseries, rtol=1e-13)
        assert_allclose(h2.results.timeseries, ref_h, rtol=TOL64)


# ------------------------------------------------------------------ pipelined bulk staging (particle chunks)
@pytest.mark.parametrize("chunk", [1, 7, 32, 1000])
def test_bulk_staging_in_particle_chunks(rand_u, chunk, monkeypatch):
    """ta_stage_bulk copies particle chunks (each with all frames) and the compute call launches one
    kernel per chunk behind it; any chunking must give the single-launch result bit for bit
    (per-particle array) and the same mean up to the summation order of the partial rows."""
    u, vel, pos, masses = rand_u
    base = VACF(u.atoms, fft=True).run()
    hbase = VH(u.atoms).run()
    hfbase = VH(u.atoms, fft=True).run()
    wbase = VACF(u.atoms, fft=False).run()
    monkeypatch.setenv("TA_B200_BULK_CHUNK", str(chunk))
    v = VACF(u.atoms, fft=True).run()
    assert np.array_equal(v.results.vacf_by_particle, base.results.vacf_by_particle)
    assert_allclose(v.results.timeseries, base.results.timeseries, rtol=1e-13, atol=1e-15)
    n_chunks = -(-u.atoms.n_atoms // chunk)
    assert v._ctx.launch_count() >= 2 * n_chunks          # K0 + K1 per chunk
    w = VACF(u.atoms, fft=False).run()
    assert np.array_equal(w.results.vacf_by_particle, wbase.results.vacf_by_particle)
    h = VH(u.atoms).run()
    assert np.array_equal(h.results.visc_by_particle, hbase.results.visc_by_particle)
    assert_allclose(h.results.timeseries, hbase.results.timeseries, rtol=1e-13)
    ref_bp, ref_ts = oracle.helfand_msd(_f64(vel), _f64(pos), masses, np.full(len(vel), np.prod(BOX[:3])), 300.0)
    assert_allclose(h.results.timeseries, ref_ts, rtol=TOL64)
    hf = VH(u.atoms, fft=True).run()
    assert np.array_equal(hf.results.visc_by_particle, hfbase.results.visc_by_particle)
    assert_allclose(hf.results.timeseries, ref_ts, rtol=TOL64)


def test_pipelined_bulk_staging_from_pinned_memory_in_a_later_context():
    """The compute call that follows ta_stage_bulk is queued behind the trajectory copies chunk by chunk.  With
    page-locked arrays those copies are truly asynchronous, so everything K1 reads must be ordered on the library's own
    streams: in round 2 the 80 KB normalisation table of the T = 10,000 plan, uploaded with a plain cudaMemcpy at
    compute time, reached the device behind the queued chunks and every chunk but the last came out as zeros -- except
    in the first context of a process, where loading the kernel image took longer than the staging.  Three contexts
    in a row, lag 0 of EVERY particle against its mean square, sampled particles against the oracle."""
    import bench

    T, N = 10000, 8288                         # four staging chunks of 2,072 particles
    vel = np.empty((T, N, 3), dtype=np.float32)
    bench.fill_random_f32(vel, seed=99)
    assert _lib.pin_array(vel)
    want0 = (_f64(vel) ** 2).sum(axis=2).mean(axis=0)
    pick = [0, 2071, 2072, 5000, 8287]
    ref_bp, _ = oracle.vacf_fft(_f64(vel[:, pick]))
    for k in range(3):
        ctx = _lib.Context([0])
        ctx.stage_begin(T, N, [0, 1, 2], np.float32, 1, None, "fp64")
        ctx.stage_bulk([vel])
        ts = ctx.vacf_fft()                    # no stage_end: queued behind the copies
        assert ctx.launch_count() >= 8         # K0 + K1 per chunk
        bp = ctx.fetch_by_particle()
        assert_allclose(bp[0], want0, rtol=1e-12, err_msg=f"context {k + 1}: lag 0 of every particle")
        assert_close_normwise(bp[:, pick], ref_bp, TOL64, f"context {k + 1}")
        assert_allclose(ts, bp.mean(axis=1), rtol=1e-12, atol=1e-13 * np.abs(ts).max())
        ctx.close()


def test_second_run_on_the_same_object_and_window(rand_u, monkeypatch):
    """run() twice on one analysis object (the context and its device buffers are reused), then a
    different frame window (buffers are rebuilt): every result matches a fresh object."""
    u, vel, _, _ = rand_u
    monkeypatch.setenv("TA_B200_BULK_CHUNK", "16")
    a = VACF(u.atoms, fft=True)
    first = np.array(a.run().results.timeseries)
    again = np.array(a.run().results.timeseries)
    assert np.array_equal(first, again)
    sub = a.run(start=10, stop=600, step=3)
    ref_bp, ref_ts = oracle.vacf_fft(_f64(vel)[10:600:3])
    assert_close_normwise(sub.results.timeseries, ref_ts, TOL64)
    assert_close_normwise(sub.results.vacf_by_particle, ref_bp, TOL64)
    # device-resident recompute after a pipelined run is a single launch and gives the same answer
    assert_allclose(a._ctx.vacf_fft(), sub.results.timeseries, rtol=1e-13, atol=1e-15)


# ------------------------------------------------------------------ opt-in FFT route of the Helfand MSD (K1 + K5 + K6)
# Comparators: the DIRECT route (fft=False, kernel K3: the reference's own sums, viscosity.py:212-226) and the oracle.
# The default route is the direct one; nothing below compares the FFT route with itself.
@pytest.mark.parametrize("dim,n_dim", [("xyz", 3), ("xz", 2), ("y", 1)])
def test_helfand_fft_route_against_exact_route(rand_u, dim, n_dim):
    """S1 - 2 S2 cancels; the lags where that costs more than the FP64 bar are re-evaluated exactly (K6), so the FFT
    route meets rtol 1e-10 on every lag of this 700-frame AR(1) trajectory, and row 0 stays exactly 0."""
    u, vel, pos, masses = rand_u
    exact = VH(u.atoms, dim_type=dim, fft=False).run()
    fast = VH(u.atoms, dim_type=dim, fft=True).run()
    assert exact.fft is False and fast.fft is True
    assert exact._ctx.fft_plan_info()["threads"] == 0 and fast._ctx.fft_plan_info()["threads"] > 0   # K1 ran only for `fast`
    assert fast._ctx.helfand_fft_refined() >= 0                                                # ... and K3 did not take over
    assert fast.results.timeseries[0] == 0.0 and np.all(fast.results.visc_by_particle[0] == 0.0)
    assert_allclose(fast.results.timeseries, exact.results.timeseries, rtol=TOL64)
    assert_allclose(fast.results.visc_by_particle[1:], exact.results.visc_by_particle[1:], rtol=TOL64)
    cols, _ = oracle.parse_dim_type(dim)
    ref_bp, ref_ts = oracle.helfand_msd(_f64(vel)[:, :, cols], _f64(pos)[:, :, cols], masses,
                                        np.full(len(vel), float(np.prod(np.float32(BOX[:3]).astype(np.float64)))), 300.0)
    assert_allclose(fast.results.timeseries, ref_ts, rtol=TOL64)
    assert_allclose(fast.results.visc_by_particle, ref_bp, rtol=TOL64)


@pytest.mark.parametrize("T,N,dim", [(10000, 310, "xyz"), (7000, 200, "xy")])
def test_helfand_fft_route_on_the_tensor_memory_kernel(T, N, dim):
    """ta_helfand_fft at lengths whose K1 is the R1 = 16 / 20 build: K1 runs WITHOUT particle sums (K5 forms its own) and
    with its output stage in tensor memory (parked residue-0 result + normalisation table).  Against the direct route on
    every (lag, particle) and against the oracle on sampled particles."""
    vel, pos = random_trajectory(T, N, seed=T + N, with_positions=True, rho=0.9)
    masses = np.random.default_rng(1).choice([1.008, 12.011, 15.999], N)
    u = make_universe(pos, vel, masses=masses, dimensions=BOX)
    exact = VH(u.atoms, dim_type=dim, fft=False).run()
    fast = VH(u.atoms, dim_type=dim, fft=True).run()
    plan = fast._ctx.fft_plan_info()
    assert plan["radices"][0] in (16, 20) and plan["tmem"] and fast._ctx.helfand_fft_refined() >= 0
    assert_allclose(fast.results.timeseries, exact.results.timeseries, rtol=TOL64)
    assert_allclose(fast.results.visc_by_particle[1:], exact.results.visc_by_particle[1:], rtol=TOL64)
    assert np.all(fast.results.visc_by_particle[0] == 0.0)
    cols, _ = oracle.parse_dim_type(dim)
    pick = [0, N // 2, N - 1]
    lags = np.array([1, 2, 17, T // 3, T // 2, T - 2, T - 1])
    vol = np.full(T, float(np.prod(np.float32(BOX[:3]).astype(np.float64))))
    ref_bp, _ = oracle.helfand_msd(_f64(vel)[:, pick][:, :, cols], _f64(pos)[:, pick][:, :, cols], masses[pick], vol, 300.0, lags=lags)
    assert_allclose(fast.results.visc_by_particle[lags][:, pick], ref_bp[lags], rtol=TOL64)


def test_helfand_default_is_the_direct_route(rand_u):
    """the O(T log T) route is opt-in; ViscosityHelfand(ag) runs the reference's direct sums (K3)."""
    u = rand_u[0]
    h = VH(u.atoms).run()
    assert h.fft is False and h._ctx.helfand_fft_refined() == 0
    assert h._ctx.fft_plan_info()["threads"] == 0                # no FFT kernel was launched


def test_helfand_fft_route_general_kernel_and_window(rand_u, monkeypatch):
    u, vel, pos, masses = rand_u
    monkeypatch.setenv("TA_B200_K1_PATH", "general")
    exact = VH(u.atoms, linear_fit_window=(20, 150), fft=False).run(start=3, stop=650, step=2)
    fast = VH(u.atoms, linear_fit_window=(20, 150), fft=True).run(start=3, stop=650, step=2)
    assert fast._ctx.fft_plan_info()["radices"][1:] != [16, 16]
    assert_allclose(fast.results.timeseries, exact.results.timeseries, rtol=TOL64)
    assert_allclose(fast.results.viscosity, exact.results.viscosity, rtol=1e-8)
    assert_allclose(fast.running_viscosity, exact.running_viscosity, rtol=TOL64)


def helfand_fft_threshold(T):
    """thr of ta_helfand_fft (csrc/ta_b200.cu): a lag is re-evaluated exactly when its un-normalised MSD < thr * sum g^2."""
    return (100.0 + T / 100.0) * 2.0 ** -53 / 2e-11


def _helfand_case(kind, T, N, seed):
    """Trajectories that stress the S1 - 2 S2 cancellation differently.  Returns float32 (vel, pos) and masses."""
    rng = np.random.default_rng(seed)
    masses = np.random.default_rng(1).choice([1.008, 12.011, 15.999], N)
    thr = helfand_fft_threshold(T)
    if kind in ("white", "spike1", "spike_all", "masses200"):
        # no correlation: MSD ~ 2 var at every lag, only the last lags are delicate
        vel = rng.standard_normal((T, N, 3)).astype(np.float32)
        pos = (10.0 * rng.standard_normal((T, N, 3))).astype(np.float32)
        if kind == "spike1":                  # one huge sample in one particle: its lags > T/2 see none of it
            vel[T // 2, 0] = 1e4
        if kind == "spike_all":               # ... in every particle: half of all lags need the exact sum
            vel[T // 2] = 1e4
        if kind == "masses200":               # light and heavy particles side by side (1 : 200)
            masses = np.where(np.arange(N) % 2 == 0, 1.0, 200.0)
    elif kind == "smooth":                    # slowly varying g: MSD << sum g^2 at short lags
        t = np.arange(T)[:, None, None]
        ph = rng.uniform(0, 6.28, (1, N, 3))
        vel = (np.sin(2e-3 * t + ph) + 2.0).astype(np.float32)
        pos = (np.cos(1.3e-3 * t + 2 * ph) + 3.0).astype(np.float32)
    elif kind == "walk":                      # positions are an unwrapped random walk (the synthetic universe of the survey)
        vel, pos = random_trajectory(T, N, seed=seed, with_positions=True, rho=0.99)
    elif kind == "ramp":                      # the reference's step trajectory: v = t, x = t^2 / 2 (g ~ t^3)
        t = np.arange(T, dtype=np.float64)[:, None, None]
        vel = np.repeat(np.repeat(t, N, 1), 3, 2).astype(np.float32)
        pos = np.repeat(np.repeat(t * t / 2, N, 1), 3, 2).astype(np.float32)
    elif kind == "frozen":                    # most particles do not move at all, a few do
        vel = np.zeros((T, N, 3), np.float32)
        pos = np.ones((T, N, 3), np.float32)
        vel[:, :3] = rng.standard_normal((T, 3, 3)).astype(np.float32)
    elif kind in ("offset_above", "offset_below"):
        # large constant + small noise: MSD[k] / sum g^2 ~ 2 a^2 (T - k) / T, placed a factor 100 above the refinement
        # threshold (only the last ~1 % of the lags are marked) or a factor 2 below it (every lag is marked)
        a = np.sqrt((100.0 if kind == "offset_above" else 0.5) * thr / 2.0)
        vel = (1.0 + a * rng.standard_normal((T, N, 3))).astype(np.float32)
        pos = np.ones((T, N, 3), np.float32)
    elif kind == "piecewise":                 # piecewise-constant moments: ten jumps of a tenth of the level
        level = 1.0 + 0.1 * rng.integers(0, 4, (11, N, 3))
        vel = np.repeat(level, -(-T // 11), axis=0)[:T].astype(np.float32)
        pos = np.ones((T, N, 3), np.float32)
    else:
        raise ValueError(kind)
    return vel, pos, masses


HELFAND_FAMILIES = [("white", 3000, 40, "xyz"), ("smooth", 3000, 40, "xyz"), ("walk", 5000, 200, "xyz"), ("ramp", 2000, 3, "xyz"),
                    ("frozen", 1600, 12, "xyz"), ("white", 10000, 160, "xyz"), ("smooth", 700, 5, "xyz"),
                    # adversarial families: spikes, offset + noise either side of the threshold,
                    # piecewise-constant, one dimension, the longest series K5 holds, masses 1 : 200
                    ("spike1", 3000, 40, "xyz"), ("spike_all", 3000, 12, "xyz"), ("offset_above", 3000, 24, "xyz"),
                    ("offset_below", 3000, 24, "xyz"), ("piecewise", 3000, 16, "xyz"), ("white", 4000, 9, "x"),
                    ("smooth", 5000, 7, "y"), ("white", 28999, 3, "xyz"), ("masses200", 5000, 64, "xyz"),
                    ("walk", 10000, 24, "xz")]


@pytest.mark.parametrize("kind,T,N,dim", HELFAND_FAMILIES)
def test_helfand_fft_route_with_exact_refinement_meets_the_fp64_bar(kind, T, N, dim):
    """ViscosityHelfand(fft=True): S1 - 2 S2 where it is good to 1e-10, the exact sum (K6) on the lags it marks, the direct
    kernel when too much is marked -- against the DIRECT route (fft=False, K3) on every (lag, particle) and against the
    oracle on sampled lags, at the FP64 bar (rtol 1e-10); the number of refined lags must be what the criterion predicts."""
    vel, pos, masses = _helfand_case(kind, T, N, seed=T + N)
    u = make_universe(pos, vel, masses=masses, dimensions=BOX)
    exact = VH(u.atoms, dim_type=dim, fft=False).run()
    fast = VH(u.atoms, dim_type=dim, fft=True).run()
    assert exact.fft is False and fast.fft is True
    refined = fast._ctx.helfand_fft_refined()
    e_bp, f_bp = np.asarray(exact.results.visc_by_particle), np.asarray(fast.results.visc_by_particle)
    assert fast.results.timeseries[0] == 0.0 and np.all(f_bp[0] == 0.0)
    assert_allclose(f_bp, e_bp, rtol=TOL64, atol=0, err_msg=f"{kind}: refined {refined} of {N * T}")
    assert_allclose(fast.results.timeseries, exact.results.timeseries, rtol=TOL64)
    # the oracle on sampled lags, both routes
    cols, D = oracle.parse_dim_type(dim)
    lags = sorted({1, 2, 3, 7, T // 10, T // 3, T // 2, T // 2 + 1, (3 * T) // 4, T - 3, T - 2, T - 1})
    vols = np.full(T, float(np.prod(np.float32(BOX[:3]).astype(np.float64))))
    ref_bp, _ = oracle.helfand_msd(_f64(vel)[:, :, cols], _f64(pos)[:, :, cols], masses, vols, 300.0, lags=lags)
    assert_allclose(f_bp[lags], ref_bp[lags], rtol=TOL64, atol=0, err_msg=f"{kind}: FFT route vs oracle")
    assert_allclose(e_bp[lags], ref_bp[lags], rtol=TOL64, atol=0, err_msg=f"{kind}: direct route vs oracle")
    # expected number of refined (particle, lag) pairs: un-normalised exact MSD below thr * sum g^2 (K5's criterion),
    # counted with the threshold 10 % lower / higher to allow for lags that sit on it
    g = masses[None, :, None] * _f64(vel)[:, :, cols] * _f64(pos)[:, :, cols]
    tot = (g ** 2).sum(axis=(0, 2))
    scale = 2 * exact.boltzmann * exact._vol_avg * exact.temp_avg * D
    msd_un = e_bp[1:] * (T - np.arange(1, T))[:, None] * scale
    lo = int((msd_un < 0.9 * helfand_fft_threshold(T) * tot[None, :]).sum())
    hi = int((msd_un < 1.1 * helfand_fft_threshold(T) * tot[None, :]).sum())
    if lo > 0.02 * N * T:
        assert refined == -1, f"{kind}: {lo} of {N * T} lags need the exact sum, the direct kernel should have taken over"
    elif hi <= 0.02 * N * T:
        assert lo <= refined <= hi, f"{kind}: refined {refined}, criterion predicts {lo}..{hi}"
    # families whose outcome is known without the count: spikes in every particle and an offset below the threshold
    # leave (almost) every lag to the exact sum -> the direct kernel; white noise and the random walk stay with the FFT
    if kind in ("spike_all", "offset_below"):
        assert refined == -1, f"{kind}: refined {refined}"
    if kind in ("white", "walk", "masses200", "offset_above", "spike1"):
        assert 0 <= refined < 0.02 * N * T, f"{kind}: refined {refined}"


def test_helfand_auto_route_falls_back_to_the_direct_sums_beyond_its_length():
    """fft='auto' on a trajectory longer than the FFT route's finishing kernel holds (T > ~29,000): the direct kernel,
    and no FFT pass is spent on finding that out."""
    T, N = 30001, 2
    vel, pos = random_trajectory(T, N, seed=3, with_positions=True, rho=0.9)
    u = make_universe(pos, vel, masses=[12.0, 16.0], dimensions=BOX)
    h = VH(u.atoms, fft="auto")
    assert h.fft is True
    h.run()
    assert h.fft is False
    assert h._ctx.launch_count() <= 4                 # K0 + K3 + the partial-row sum: no K1
    lags = [1, 2, 15000, 30000]
    ref_bp, _ = oracle.helfand_msd(_f64(vel), _f64(pos), np.array([12.0, 16.0]),
                                   np.full(T, float(np.prod(np.float32(BOX[:3]).astype(np.float64)))), 300.0, lags=lags)
    assert_allclose(np.asarray(h.results.visc_by_particle)[lags], ref_bp[lags], rtol=TOL64)


# ------------------------------------------------------------------ FP32 mode (stated tolerance 1e-5)
def test_fp32_mode_helfand(rand_u):
    u, vel, pos, masses = rand_u
    vols = np.full(700, float(np.prod(np.float32(BOX[:3]).astype(np.float64))))
    for dim in ("xyz", "y"):
        cols, _ = oracle.parse_dim_type(dim)
        ref_bp, ref_ts = oracle.helfand_msd(_f64(vel)[:, :, cols], _f64(pos)[:, :, cols], masses, vols, 300.0)
        h = VH(u.atoms, dim_type=dim, precision="fp32").run()
        assert h.results.timeseries[0] == 0.0
        assert_close_normwise(h.results.timeseries, ref_ts, TOL32, "fp32 Helfand timeseries")
        assert_close_normwise(h.results.visc_by_particle, ref_bp, TOL32, "fp32 Helfand by particle")
    with pytest.raises(ValueError, match="fp64"):
        VH(u.atoms, precision="fp32", fft=True)


def test_fp32_mode_at_full_length():
    """FP32 mode at T = 10,000 (BASELINE configs[3] length), where single-precision accumulation error is largest:
    FFT route (the three-pass FP32 kernel), windowed route and Helfand against the FP64 oracle at 1e-5 normwise."""
    T, N = 10000, 12
    vel, pos = random_trajectory(T, N, seed=77, with_positions=True, rho=0.9)
    masses = np.random.default_rng(5).choice([1.008, 12.011, 15.999], N)
    u = make_universe(pos, vel, masses=masses, dimensions=BOX)
    ref_bp, ref_ts = oracle.vacf_fft(_f64(vel))
    for fft in (True, False):
        v = VACF(u.atoms, fft=fft, precision="fp32").run()
        if fft:
            assert v._ctx.fft_plan_info()["radices"] == [20, 16, 16]
        # FFT route: the float rounding floor of the un-normalised correlation is divided by the number of origins, so
        # the stated 1e-5 covers the lags with at least 8 origins and the last 8 are within 1e-4 (DESIGN.md "FP32 mode")
        cut = -8 if fft else None
        assert_close_normwise(v.results.timeseries[:cut], ref_ts[:cut], TOL32, f"fp32 fft={fft} timeseries")
        assert_close_normwise(v.results.vacf_by_particle[:cut], ref_bp[:cut], TOL32, f"fp32 fft={fft} by particle")
        assert_close_normwise(v.results.vacf_by_particle, ref_bp, 1e-4, f"fp32 fft={fft} by particle, last lags")
    lags = [1, 2, 100, 2500, 5000, 7500, 9990, 9999]
    vols = np.full(T, float(np.prod(np.float32(BOX[:3]).astype(np.float64))))
    ref_h, _ = oracle.helfand_msd(_f64(vel), _f64(pos), masses, vols, 300.0, lags=lags)
    h = VH(u.atoms, precision="fp32").run()
    hb = np.asarray(h.results.visc_by_particle)
    assert_close_normwise(hb[lags], ref_h[lags], TOL32, "fp32 Helfand by particle")


# ------------------------------------------------------------------ series longer than shared memory (direct routes)
def test_windowed_routes_beyond_shared_memory():
    """T = 16,000 frames: series + lag sums of one particle (17 T bytes) exceed the 227 KB of an SM, so K2/K3
    keep them in a per-CTA global scratch area instead; same arithmetic, same parity bar."""
    T, N = 16000, 2
    vel, pos = random_trajectory(T, N, seed=21, with_positions=True, rho=0.95)
    masses = np.array([12.011, 1.008])
    u = make_universe(pos, vel, masses=masses, dimensions=BOX)
    w = VACF(u.atoms, dim_type="x", fft=False).run()
    ref_bp, ref_ts = oracle.vacf_windowed(_f64(vel)[:, :, [0]])
    assert_close_normwise(w.results.vacf_by_particle, ref_bp, TOL64)
    assert_close_normwise(w.results.timeseries, ref_ts, TOL64)
    h = VH(u.atoms, dim_type="x").run()
    _, ref_h = oracle.helfand_msd(_f64(vel)[:, :, [0]], _f64(pos)[:, :, [0]], masses, np.full(T, np.prod(BOX[:3])), 300.0)
    assert_allclose(h.results.timeseries, ref_h, rtol=TOL64)
