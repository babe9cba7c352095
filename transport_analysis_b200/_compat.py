"""The drop-in boundary: MDAnalysis' ``AnalysisBase`` protocol.

The reference classes subclass ``MDAnalysis.analysis.base.AnalysisBase``
(transport_analysis/velocityautocorr.py:52-54,72,120; viscosity.py:16-19,26,89).
When MDAnalysis is importable the real classes are used unchanged.  It is not
installed in the build image, so a protocol-compatible stand-in is provided:
``run()`` template-method driver, ``Results``, ``NoDataError``, the
``UpdatingAtomGroup`` sentinel, ``units.constants`` and a MemoryReader-backed
synthetic ``Universe`` (the reference's own recipe for in-memory universes,
transport_analysis/tests/utils.py:8-77).  The stand-in is harness for the
boundary -- it contains no part of the hot path.
"""
from __future__ import annotations

import math

import numpy as np

try:  # pragma: no cover - exercised only where MDAnalysis is installed
    from MDAnalysis.analysis.base import AnalysisBase, Results  # type: ignore
    from MDAnalysis.core.groups import UpdatingAtomGroup  # type: ignore
    from MDAnalysis.exceptions import NoDataError  # type: ignore
    from MDAnalysis.units import constants  # type: ignore

    HAVE_MDANALYSIS = True
except ImportError:
    HAVE_MDANALYSIS = False

    class NoDataError(ValueError, AttributeError):
        """Raised when the trajectory lacks the requested data."""

    # MDAnalysis.units.constants (kJ/(mol K) for the Boltzmann constant)
    constants = {
        "Boltzmann_constant": 8.314462159e-3,
        "N_Avogadro": 6.02214129e23,
    }

    class Results(dict):
        """Attribute-access dictionary (MDAnalysis.analysis.base.Results)."""

        def __getattr__(self, key):
            try:
                return self[key]
            except KeyError as exc:
                raise AttributeError(f"'Results' object has no attribute '{key}'") from exc

        def __setattr__(self, key, value):
            self[key] = value

        def __delattr__(self, key):
            try:
                del self[key]
            except KeyError as exc:
                raise AttributeError(key) from exc

    class AnalysisBase:
        """Template-method driver: ``run()`` = ``_setup_frames`` + ``_prepare``
        + ``_single_frame`` per frame + ``_conclude``."""

        def __init__(self, trajectory, verbose=False, **kwargs):
            self._trajectory = trajectory
            self._verbose = verbose
            self.results = Results()

        def _setup_frames(self, trajectory, start=None, stop=None, step=None, frames=None):
            self._trajectory = trajectory
            if frames is not None:
                if not all(opt is None for opt in (start, stop, step)):
                    raise ValueError("start/stop/step cannot be combined with frames")
                frames = np.asarray(frames)
                self._sliced_trajectory = trajectory[frames]
                self.start = self.stop = self.step = None
                self.n_frames = len(frames)
            else:
                start, stop, step = trajectory.check_slice_indices(start, stop, step)
                self.start, self.stop, self.step = start, stop, step
                self._sliced_trajectory = trajectory[start:stop:step]
                self.n_frames = len(range(start, stop, step))
            self.frames = np.zeros(self.n_frames, dtype=int)
            self.times = np.zeros(self.n_frames)

        def _prepare(self):
            pass

        def _single_frame(self):
            raise NotImplementedError("Only implemented in child classes")

        def _conclude(self):
            pass

        def run(self, start=None, stop=None, step=None, frames=None, verbose=None, **kwargs):
            self._setup_frames(self._trajectory, start=start, stop=stop, step=step, frames=frames)
            self._prepare()
            for i, ts in enumerate(self._sliced_trajectory):
                self._frame_index = i
                self._ts = ts
                self.frames[i] = ts.frame
                self.times[i] = ts.time
                self._single_frame()
            self._conclude()
            return self

    # ------------------------------------------------------------------
    # in-memory universe
    # ------------------------------------------------------------------
    def _box_volume(dimensions):
        if dimensions is None:
            return 0.0
        a, b, c, al, be, ga = (float(v) for v in dimensions)
        if a == 0 or b == 0 or c == 0:
            return 0.0
        if al == 90.0 and be == 90.0 and ga == 90.0:
            return a * b * c
        ca, cb, cg = (math.cos(math.radians(v)) for v in (al, be, ga))
        return a * b * c * math.sqrt(max(0.0, 1 - ca * ca - cb * cb - cg * cg + 2 * ca * cb * cg))

    class Timestep:
        def __init__(self, reader, frame):
            self._reader = reader
            self.frame = int(frame)

        @property
        def time(self):
            return self.frame * self._reader.dt

        @property
        def dt(self):
            return self._reader.dt

        @property
        def has_positions(self):
            return self._reader.coordinate_array is not None

        @property
        def has_velocities(self):
            return self._reader.velocity_array is not None

        @property
        def positions(self):
            if not self.has_positions:
                raise NoDataError("This Timestep has no position information")
            return self._reader.coordinate_array[self.frame]

        @property
        def velocities(self):
            if not self.has_velocities:
                raise NoDataError("This Timestep has no velocities information")
            return self._reader.velocity_array[self.frame]

        @property
        def dimensions(self):
            d = self._reader.dimensions_array
            return None if d is None else d[self.frame]

        @property
        def volume(self):
            return _box_volume(self.dimensions)

    class _SlicedTrajectory:
        def __init__(self, reader, frames):
            self._reader, self._frames = reader, frames

        def __len__(self):
            return len(self._frames)

        def __iter__(self):
            for f in self._frames:
                yield self._reader._goto(f)

    class MemoryReader:
        """Trajectory held in memory as float32 ``[frames, atoms, 3]`` arrays
        (MDAnalysis.coordinates.memory.MemoryReader, 'fac' order)."""

        stored_order = "fac"

        def __init__(self, coordinate_array=None, velocities=None, dimensions=None, dt=1.0):
            def f32(a):
                return None if a is None else np.ascontiguousarray(a, dtype=np.float32)

            self.coordinate_array = f32(coordinate_array)
            self.velocity_array = f32(velocities)
            ref = self.coordinate_array if self.coordinate_array is not None else self.velocity_array
            if ref is None:
                raise ValueError("MemoryReader needs positions or velocities")
            self.n_frames, self.n_atoms = ref.shape[0], ref.shape[1]
            if dimensions is None:
                self.dimensions_array = None
            else:
                d = np.asarray(dimensions, dtype=np.float32)
                if d.ndim == 1:
                    d = np.tile(d, (self.n_frames, 1))
                self.dimensions_array = d
            self.dt = float(dt)
            self.ts = Timestep(self, 0)

        def get_array(self):
            return self.coordinate_array

        def __len__(self):
            return self.n_frames

        def _goto(self, frame):
            self.ts = Timestep(self, frame)
            return self.ts

        def __iter__(self):
            for f in range(self.n_frames):
                yield self._goto(f)

        def __getitem__(self, item):
            if isinstance(item, (int, np.integer)):
                f = int(item)
                if f < 0:
                    f += self.n_frames
                if not 0 <= f < self.n_frames:
                    raise IndexError(f"frame {item} out of range")
                return self._goto(f)
            if isinstance(item, slice):
                start, stop, step = self.check_slice_indices(item.start, item.stop, item.step)
                return _SlicedTrajectory(self, range(start, stop, step))
            return _SlicedTrajectory(self, [int(f) for f in np.asarray(item)])

        def check_slice_indices(self, start, stop, step):
            """Same contract as ProtoReader.check_slice_indices for step > 0."""
            for v in (start, stop, step):
                if v is not None and not isinstance(v, (int, np.integer)):
                    raise TypeError("Slice indices are not integers")
            if step == 0:
                raise ValueError("Step size is zero")
            n = len(self)
            step = 1 if step is None else int(step)
            if step > 0:
                start = 0 if start is None else int(start)
                stop = n if stop is None else int(stop)
                if start < 0:
                    start += n
                if stop < 0:
                    stop += n
                start = min(max(start, 0), n)
                stop = min(max(stop, 0), n)
                if stop < start:
                    stop = start
            else:
                start, stop, _ = slice(start, stop, step).indices(n)
            return start, stop, step

    class AtomGroup:
        def __init__(self, universe, ix):
            self.universe = universe
            self.ix = np.asarray(ix, dtype=np.int64)

        def __len__(self):
            return len(self.ix)

        @property
        def n_atoms(self):
            return len(self.ix)

        def __getitem__(self, item):
            return AtomGroup(self.universe, np.atleast_1d(self.ix[item]))

        @property
        def positions(self):
            return self.universe.trajectory.ts.positions[self.ix]

        @property
        def velocities(self):
            return self.universe.trajectory.ts.velocities[self.ix]

        @property
        def masses(self):
            if self.universe._masses is None:
                raise NoDataError("This Universe does not contain masses")
            return self.universe._masses[self.ix].astype(np.float64)

    class UpdatingAtomGroup(AtomGroup):
        """Sentinel for selections that are re-evaluated every frame."""

    class Universe:
        """Minimal in-memory universe: ``Universe(positions=..., velocities=...,
        masses=..., dimensions=..., dt=...)`` with float ``[frames, atoms, 3]``
        arrays (either may be None)."""

        def __init__(self, positions=None, velocities=None, masses=None, dimensions=None, dt=1.0):
            self.trajectory = MemoryReader(positions, velocities=velocities, dimensions=dimensions, dt=dt)
            n = self.trajectory.n_atoms
            self._masses = None if masses is None else np.asarray(masses, dtype=np.float64).reshape(n)
            self.atoms = AtomGroup(self, np.arange(n))

        def add_TopologyAttr(self, name, values):
            if name != "masses":
                raise NotImplementedError(name)
            self._masses = np.asarray(values, dtype=np.float64).reshape(len(self.atoms))

        def select_atoms_updating(self, ix):
            return UpdatingAtomGroup(self, ix)
