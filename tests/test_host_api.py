"""Host-side behaviour that needs no GPU: the C-ABI library loads and exports
every symbol include/ta_b200.h declares; constructor / run() error behaviour
matches the reference (tests/test_velocityautocorr.py:132-149,201-217;
tests/test_viscosity.py:139-155); the AnalysisBase stand-in slices frames like
MDAnalysis; the product path fails loudly without a device."""
import os
import re

import numpy as np
import pytest

import transport_analysis_b200 as tab
from transport_analysis_b200 import _compat, _lib
from transport_analysis_b200.synthetic import make_universe, random_trajectory
from transport_analysis_b200.velocityautocorr import VelocityAutocorr as VACF
from transport_analysis_b200.viscosity import ViscosityHelfand as VH

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def u():
    vel, pos = random_trajectory(12, 10, seed=1, with_positions=True)
    return make_universe(pos, vel, masses=np.full(10, 15.999), dimensions=[20, 20, 20, 90, 90, 90])


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "ta_b200.h")).read()
    declared = set(re.findall(r"\b(ta_[a-z0-9_]+)\s*\(", header))
    declared.discard("ta_ctx")
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    lib = _lib.load_library()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.ta_version() >= 100


def test_no_oracle_import_in_product():
    pkg = os.path.join(ROOT, "transport_analysis_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_fails_loudly_without_gpu(u):
    if _lib.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(tab.BackendError, match="no CPU fallback"):
        VACF(u.atoms).run()
    with pytest.raises(tab.BackendError, match="no CPU fallback"):
        VH(u.atoms).run()


def test_no_velocities():
    _, pos = random_trajectory(5, 10, with_positions=True)
    u_no_vels = make_universe(pos, None, masses=np.ones(10), dimensions=[2, 2, 2, 90, 90, 90])
    with pytest.raises(_compat.NoDataError, match="VACF computation requires velocities"):
        VACF(u_no_vels.atoms, fft=False).run()
    with pytest.raises(_compat.NoDataError, match="Helfand viscosity computation requires"):
        VH(u_no_vels.atoms).run()


def test_no_box_volume():
    vel, pos = random_trajectory(5, 4, with_positions=True)
    u_nobox = make_universe(pos, vel, masses=np.ones(4))
    with pytest.raises(_compat.NoDataError, match="Helfand viscosity computation requires"):
        VH(u_nobox.atoms).run()


def test_updating_ag_rejected(u):
    if _compat.HAVE_MDANALYSIS:
        pytest.skip("needs the stand-in universe")
    upd = u.select_atoms_updating([0, 1, 2])
    with pytest.raises(TypeError, match="UpdatingAtomGroups are not valid"):
        VACF(upd, fft=False)
    with pytest.raises(TypeError, match="UpdatingAtomGroups are not valid"):
        VH(upd)


@pytest.mark.parametrize("dimtype", ["foo", "bar", "yx", "zyx"])
def test_dimtype_error(u, dimtype):
    with pytest.raises(ValueError, match=f"invalid dim_type: {dimtype}"):
        VACF(u.atoms, dim_type=dimtype)
    with pytest.raises(ValueError, match=f"invalid dim_type: {dimtype}"):
        VH(u.atoms, dim_type=dimtype)


def test_dimtype_case_insensitive(u):
    v = VACF(u.atoms, dim_type="XZ")
    assert v.dim_type == "xz" and v._dim == [0, 2] and v.dim_fac == 2


def test_helpers_before_run(u):
    v = VACF(u.atoms, fft=False)
    for fn in (v.plot_vacf, v.self_diffusivity_gk, v.self_diffusivity_gk_odd, v.plot_running_integral):
        with pytest.raises(RuntimeError, match="Analysis must be run"):
            fn()


@pytest.mark.skipif(_compat.HAVE_MDANALYSIS, reason="tests the stand-in driver")
def test_standin_frame_slicing_matches_range():
    vel, _ = random_trajectory(50, 2)
    uu = make_universe(None, vel)

    class Probe(_compat.AnalysisBase):
        def __init__(self, traj):
            super().__init__(traj)

        def _prepare(self):
            self.seen = []

        def _single_frame(self):
            self.seen.append((self._frame_index, self._ts.frame))

    for sl in [(None, None, None), (10, 40, 7), (3, None, 2), (None, 20, None), (-10, None, 3)]:
        p = Probe(uu.trajectory).run(*sl)
        expect = list(range(50))[slice(*sl)]
        assert [f for _, f in p.seen] == expect
        assert p.n_frames == len(expect)
        assert list(p.frames) == expect
        np.testing.assert_array_equal(p.times, np.array(expect, dtype=float))


def test_running_viscosity_reproduces_the_demo_notebook(u):
    """docs/tutorials/viscosity_early_demo.ipynb: timeseries (:47-48) and the running viscosity it
    prints (:133-134) for a 10-frame run whose frame times are 1..10."""
    ts = np.array([0.0, 96.43610866, 94.54194053, 86.22211806, 85.18650814, 93.90772559, 108.7836928, 97.7115102,
                   96.4890803, 139.72614008])
    want = np.array([48.21805433, 31.51398018, 21.55552951, 17.03730163, 15.6512876, 15.54052754, 12.21393878,
                     10.72100892, 13.97261401])
    h = VH(u.atoms)
    with pytest.raises(RuntimeError, match="must be run"):
        h.running_viscosity
    h.results.timeseries = ts
    h.times = np.arange(1.0, 11.0)
    h.n_frames = 10
    np.testing.assert_allclose(h.running_viscosity, want, rtol=1e-8)


def test_helfand_fft_route_needs_fp64(u):
    with pytest.raises(ValueError, match="fp64"):
        VH(u.atoms, fft=True, precision="fp32")


def test_helfand_route_argument(u):
    """fft=False (default): the reference's direct lag sums.  True: the O(T log T) route with exact
    refinement (FP64 only).  'auto': that route where it applies, the direct sums in the FP32 mode.  Anything else is
    refused."""
    assert VH(u.atoms).fft is False and not VH(u.atoms)._fft_auto
    assert VH(u.atoms, precision="fp32").fft is False
    assert VH(u.atoms, fft="auto").fft is True and VH(u.atoms, fft="auto")._fft_auto
    assert VH(u.atoms, fft="auto", precision="fp32").fft is False
    assert VH(u.atoms, fft=True).fft is True and not VH(u.atoms, fft=True)._fft_auto
    for bad in ("yes", 2, None):
        with pytest.raises(ValueError, match="fft must be"):
            VH(u.atoms, fft=bad)


def test_unsupported_error_is_a_backend_error():
    from transport_analysis_b200 import _lib

    assert issubclass(_lib.UnsupportedError, _lib.BackendError) and _lib.TA_ERR_UNSUPPORTED == -4
    with open(os.path.join(ROOT, "include", "ta_b200.h")) as f:
        assert "#define TA_ERR_UNSUPPORTED (-4)" in f.read()


def test_lazy_by_particle_behaves_like_the_array():
    """A per-particle result left on the GPU is fetched on first use and then acts as the (n_frames, n_particles) array the
    reference stores: numpy functions, ndarray methods, arithmetic, indexing; particle ranges without the full fetch; a handle
    of an earlier run that was never read refuses to return the next run's values."""
    from transport_analysis_b200._staging import LazyByParticle

    class FakeCtx:
        def __init__(self):
            self.data = np.arange(12.0).reshape(3, 4)      # (T, N)
            self.fetches = []

        def fetch_by_particle(self, atom0=0, natoms=None):
            self.fetches.append((atom0, natoms))
            return self.data[:, atom0:atom0 + natoms].copy()

    ctx = FakeCtx()
    lz = LazyByParticle(ctx, 3, 4)
    assert lz.shape == (3, 4) and lz.ndim == 2 and len(lz) == 3 and lz.size == 12 and lz.nbytes == 96 and not ctx.fetches
    np.testing.assert_array_equal(lz.particles(1, 3), ctx.data[:, 1:3])
    assert ctx.fetches == [(1, 2)]
    np.testing.assert_array_equal(lz.mean(axis=1), ctx.data.mean(axis=1))        # ndarray method -> one full fetch
    assert ctx.fetches == [(1, 2), (0, 4)]
    np.testing.assert_array_equal(np.mean(lz, axis=1), ctx.data.mean(axis=1))
    np.testing.assert_array_equal(lz * 2 + 1, ctx.data * 2 + 1)
    np.testing.assert_array_equal(1 - lz, 1 - ctx.data)
    np.testing.assert_array_equal(lz[1:, ::2], ctx.data[1:, ::2])
    np.testing.assert_array_equal(lz.T, ctx.data.T)
    np.testing.assert_array_equal(lz.particles(0, 2), ctx.data[:, :2])
    assert len(ctx.fetches) == 2                                                  # everything after the first use is cached
    lz.invalidate()                                                               # already read: stays valid
    np.testing.assert_array_equal(np.asarray(lz), ctx.data)
    unread = LazyByParticle(ctx, 3, 4)
    unread.invalidate()
    with pytest.raises(RuntimeError, match="later run"):
        np.asarray(unread)
    with pytest.raises(RuntimeError, match="later run"):
        unread.particles(0, 1)
    assert "stale" in repr(unread) and "fetched" in repr(lz)
