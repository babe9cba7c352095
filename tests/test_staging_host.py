"""Host side of `_single_frame` staging and of the analysis classes, on the CPU with a recording stand-in for the backend
context (no arithmetic: it only notes which C-ABI calls the host code makes and hands back canned results).

Covers what the GPU tests cannot look at from outside: how frames are packed into the pinned slabs and committed, when the
whole-trajectory fast path is taken and when it must not be, what `_conclude` asks the backend for, and the Green-Kubo
helpers of `VelocityAutocorr` (reference velocityautocorr.py:287-422) on a known timeseries."""
import numpy as np
import pytest
from scipy import integrate

from transport_analysis_b200 import _staging
from transport_analysis_b200._staging import FrameStager
from transport_analysis_b200.synthetic import make_universe
from transport_analysis_b200.velocityautocorr import VelocityAutocorr as VACF
from transport_analysis_b200.viscosity import ViscosityHelfand as VH


class FakeContext:
    """Records the calls; slabs hold `cap` frames.  Stands in for transport_analysis_b200._lib.Context (same
    constructor), so that `isinstance(x, _lib.Context)` keeps working in the host code."""

    created = []

    def __init__(self, devices=None, cap=4, **kwargs):
        self.cap, self.calls, self.slabs = cap, [], []
        self.T = self.N = None
        FakeContext.created.append(self)

    def stage_begin(self, T, N, dims, dtype, n_fields, masses, precision):
        self.T, self.N, self.n_fields, self.dtype = T, N, n_fields, dtype
        self.calls.append(("begin", T, N, list(dims), np.dtype(dtype).name, n_fields, precision))

    def stage_slot(self):
        slab = np.full((self.cap, self.n_fields, self.N, 3), np.nan, dtype=self.dtype)
        self.slabs.append(slab)
        return slab

    def stage_commit(self, frame0, nframes):
        self.calls.append(("commit", frame0, nframes, self.slabs[-1][:nframes].copy()))

    def stage_bulk(self, fields, atom_first, frame_first, frame_step, nframes):
        self.calls.append(("bulk", [f.shape for f in fields], atom_first, frame_first, frame_step, nframes))

    def stage_end(self):
        self.calls.append(("end",))

    def vacf_fft(self):
        self.calls.append(("vacf_fft",))
        return np.exp(-np.arange(self.T) / 5.0)

    def vacf_windowed(self):
        self.calls.append(("vacf_windowed",))
        return np.exp(-np.arange(self.T) / 5.0)

    def helfand(self, volumes, boltzmann, temp_avg, fft=False):
        self.calls.append(("helfand", np.array(volumes), boltzmann, temp_avg, fft))
        return np.arange(self.T, dtype=np.float64) * 2.0

    def fetch_by_particle(self, atom0=0, natoms=None):
        natoms = self.N - atom0 if natoms is None else natoms
        self.calls.append(("fetch", atom0, natoms))
        return np.zeros((self.T, natoms))


class Recorder:
    """The context the code under test created (exactly one per analysis object)."""

    @property
    def ctx(self):
        assert len(FakeContext.created) == 1, "one backend context per run"
        return FakeContext.created[0]

    @property
    def calls(self):
        return self.ctx.calls


@pytest.fixture
def fake(monkeypatch):
    FakeContext.created = []
    monkeypatch.setattr(_staging._lib, "Context", FakeContext)
    return Recorder()


def names(rec):
    return [c[0] for c in rec.calls]


def traj(T=10, N=6, seed=0):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((T, N, 3)).astype(np.float32), rng.standard_normal((T, N, 3)).astype(np.float32)


def test_per_frame_path_packs_slabs_and_commits_the_tail(fake):
    vel, _ = traj(T=10)
    st = FrameStager([0], 10, 6, [0, 1, 2], 1, None, "fp64")
    for i in range(10):
        st.add_frame(i, vel[i])
    st.finish()
    commits = [c for c in fake.calls if c[0] == "commit"]
    assert [(c[1], c[2]) for c in commits] == [(0, 4), (4, 4), (8, 2)]             # full slabs, then the tail
    np.testing.assert_array_equal(np.concatenate([c[3][:, 0] for c in commits]), vel)
    assert names(fake)[0] == "begin" and names(fake)[-1] == "end"
    assert fake.calls[0][4] == "float32"                                            # the source dtype travels as it is


def test_float64_sources_are_staged_as_float64(fake):
    st = FrameStager([0], 2, 3, [0], 1, None, "fp64")
    st.add_frame(0, np.ones((3, 3)))
    assert fake.calls[0][4] == "float64"


def test_bulk_path_taken_for_a_regular_window_of_contiguous_atoms(fake):
    vel, pos = traj(T=20, N=8)
    u = make_universe(pos, vel, masses=np.ones(8), dimensions=[10, 10, 10, 90, 90, 90])
    a = VACF(u.atoms[2:7], fft=True).run(start=3, stop=18, step=5)
    assert ("bulk", [(20, 8, 3)], 2, 3, 5, 3) in fake.calls and "commit" not in names(fake)
    assert names(fake)[-2:] == ["vacf_fft", "fetch"] and a.n_frames == 3
    assert "end" not in names(fake)                                                 # the compute call is queued behind the copy


@pytest.mark.parametrize("pick", [[0, 2, 3], [3, 2, 1]])
def test_bulk_path_refused_for_scattered_atoms(fake, pick):
    vel, pos = traj(T=6, N=5)
    u = make_universe(pos, vel)
    VACF(u.atoms[pick], fft=False).run()
    assert "bulk" not in names(fake) and names(fake).count("commit") == 2 and "vacf_windowed" in names(fake)
    first = [c for c in fake.calls if c[0] == "commit"][0]
    np.testing.assert_array_equal(first[3][:, 0], vel[:4][:, pick])                 # gathered per frame, in group order


def test_helfand_asks_for_both_fields_masses_and_volumes(fake):
    vel, pos = traj(T=7, N=4)
    masses = np.array([1.0, 12.0, 16.0, 14.0])
    u = make_universe(pos, vel, masses=masses, dimensions=[2, 3, 4, 90, 90, 90])
    h = VH(u.atoms, temp_avg=250.0, dim_type="XZ", linear_fit_window=(1, 6)).run()
    begin = fake.calls[0]
    assert begin[0] == "begin" and begin[3] == [0, 2] and begin[5] == 2
    bulk = [c for c in fake.calls if c[0] == "bulk"][0]
    assert bulk[1] == [(7, 4, 3), (7, 4, 3)]                                        # velocities, then positions
    call = [c for c in fake.calls if c[0] == "helfand"][0]
    np.testing.assert_allclose(call[1], np.full(7, 24.0))
    assert call[3] == 250.0 and call[4] is False                                    # default route: the direct lag sums
    # the linear fit of the reference (viscosity.py:235-245): x starts at lag 1, y at lag 0 -> slope of y = 2 k is 2
    assert h.results.viscosity == pytest.approx(2.0)


def test_green_kubo_helpers_follow_the_reference(fake):
    vel, _ = traj(T=40, N=3)
    u = make_universe(None, vel)
    a = VACF(u.atoms, dim_type="xy", fft=True).run()
    ts, t = a.results.timeseries, a.times
    assert a.self_diffusivity_gk() == pytest.approx(integrate.trapezoid(ts, t) / 2)            # /dim_fac (:316-322)
    assert a.self_diffusivity_gk(start=2, stop=30, step=3) == pytest.approx(integrate.trapezoid(ts[2:30:3], t[2:30:3]) / 2)
    assert a.self_diffusivity_gk_odd(stop=39) == pytest.approx(integrate.simpson(y=ts[:39], x=t[:39]) / 2)
    with pytest.raises(RuntimeError):
        VACF(u.atoms).self_diffusivity_gk()


# ------------------------------------------------------------------ when the whole-trajectory path may be taken
def test_bulk_path_from_an_explicit_regular_frame_list(fake):
    """MDAnalysis >= 2.8 hands _setup_frames an explicit frame list (start / stop / step are then None): a regular list is
    still streamed as a whole, an irregular one goes frame by frame."""
    vel, _ = traj(T=20, N=4)
    u = make_universe(None, vel)
    a = VACF(u.atoms, fft=True).run(frames=[3, 8, 13])
    assert a.start is None and ("bulk", [(20, 4, 3)], 0, 3, 5, 3) in fake.calls and "commit" not in names(fake)


def test_per_frame_path_for_an_irregular_frame_list(fake):
    vel, _ = traj(T=20, N=4)
    u = make_universe(None, vel)
    VACF(u.atoms, fft=True).run(frames=[3, 4, 9])
    assert "bulk" not in names(fake)
    commits = [c for c in fake.calls if c[0] == "commit"]
    np.testing.assert_array_equal(np.concatenate([c[3][:, 0] for c in commits]), vel[[3, 4, 9]])


def test_no_bulk_path_under_on_the_fly_transformations_or_on_request(fake):
    """Transformations act on Timesteps, which the whole-trajectory path never builds; staging='per_frame' asks for the
    _single_frame path outright."""
    vel, _ = traj(T=6, N=3)
    u = make_universe(None, vel)
    u.trajectory.transformations = (lambda ts: ts,)
    VACF(u.atoms, fft=True).run()
    assert "bulk" not in names(fake) and names(fake).count("commit") == 2
    FakeContext.created = []
    u2 = make_universe(None, vel)
    VACF(u2.atoms, fft=True, staging="per_frame").run()
    assert "bulk" not in names(fake)
    with pytest.raises(ValueError, match="staging"):
        VACF(u2.atoms, staging="sometimes")


def test_regular_frame_window_and_gather_index():
    class A:
        pass

    a = A()
    a.n_frames, a.start, a.step = 4, 2, 3
    assert _staging.regular_frame_window(a) == (2, 3)
    a.step = -1
    assert _staging.regular_frame_window(a) is None
    a.start = a.step = None
    a._sliced_trajectory = type("S", (), {"_frames": [5, 7, 9, 11]})()
    assert _staging.regular_frame_window(a) == (5, 2)
    a._sliced_trajectory = type("S", (), {"frames": np.array([5, 7, 10, 11])})()
    assert _staging.regular_frame_window(a) is None
    a._sliced_trajectory = None
    assert _staging.regular_frame_window(a) is None
    assert _staging.gather_index(np.arange(3, 9)) == slice(3, 9)
    ix = _staging.gather_index([4, 2, 7])
    assert isinstance(ix, np.ndarray) and list(ix) == [4, 2, 7]
    src, out = np.arange(30.0).reshape(10, 3), np.empty((3, 3))
    _staging._gather(src, ix, out)
    np.testing.assert_array_equal(out, src[[4, 2, 7]])
    out6 = np.empty((6, 3))
    _staging._gather(src, slice(3, 9), out6)
    np.testing.assert_array_equal(out6, src[3:9])


def test_against_a_real_mdanalysis_memory_reader(fake):
    """With MDAnalysis installed the same classes sit on the real AnalysisBase / MemoryReader (not available in this image)."""
    pytest.importorskip("MDAnalysis")
    vel, pos = traj(T=12, N=5)
    u = make_universe(pos, vel, masses=np.ones(5), dimensions=[10, 10, 10, 90, 90, 90])
    VACF(u.atoms, fft=True).run(start=2, stop=11, step=3)
    assert any(c[0] == "bulk" and c[3:] == (2, 3, 3) for c in fake.calls) or "commit" in names(fake)
