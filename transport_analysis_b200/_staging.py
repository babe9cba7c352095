"""Host side of `_single_frame` staging, shared by both analysis classes.

The reference copies one frame at a time into a host ``[T, N, D]`` float64
array (velocityautocorr.py:150-152,192-194; viscosity.py:128-134,189-199).
Here frames are batched into pinned slabs owned by the backend and streamed
to HBM with asynchronous copies that overlap the trajectory loop; when the
reader already holds the whole trajectory in memory (MemoryReader) the
per-frame loop is bypassed and the array is streamed directly.
"""
from __future__ import annotations

import os

import numpy as np

from . import _lib


def resolve_devices(devices):
    if isinstance(devices, _lib.Context):
        return devices          # an open context to run on (bench.py's multi-rank mode)
    if devices is None:
        env = os.environ.get("TA_B200_DEVICES")
        if env:
            return [int(x) for x in env.split(",") if x.strip() != ""]
        return [0]
    if isinstance(devices, (int, np.integer)):
        return [int(devices)]
    return [int(d) for d in devices]


class LazyByParticle(np.lib.mixins.NDArrayOperatorsMixin):
    """Per-particle result that is still on the GPUs, behaving like the ``(n_frames, n_particles)`` float64 array
    the reference stores (velocityautocorr.py:145-147, viscosity.py:117-119).

    The array is fetched the first time its values are needed -- ``np.asarray(handle)``, ``handle[...]``, arithmetic,
    any numpy function or ndarray method / attribute (``handle.mean(axis=1)``, ``handle.T`` ...) -- and kept;
    ``handle.particles(a, b)`` fetches only particles a..b-1 without materialising the rest.  A later ``run()`` on the
    same analysis object overwrites the device buffer: a handle of the earlier run that was never read then refuses
    to produce values instead of returning the new run's.
    """

    def __init__(self, ctx: "_lib.Context", T: int, N: int):
        self._ctx, self.shape, self.dtype, self.ndim = ctx, (T, N), np.dtype(np.float64), 2
        self._cache = None
        self._stale = False

    def _check_fresh(self):
        if self._stale:
            raise RuntimeError("this per-particle result was left on the GPU and a later run() has overwritten it; "
                               "read it (np.asarray) before running again")

    def invalidate(self):
        """Called by the analysis classes when a new run() reuses the device buffers."""
        if self._cache is None:
            self._stale = True

    def particles(self, start: int, stop: int) -> np.ndarray:
        if self._cache is not None:
            return self._cache[:, start:stop]
        self._check_fresh()
        return self._ctx.fetch_by_particle(start, stop - start)

    def __array__(self, dtype=None, copy=None):
        if self._cache is None:
            self._check_fresh()
            self._cache = self._ctx.fetch_by_particle(0, self.shape[1])
        return self._cache if dtype is None else self._cache.astype(dtype)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        conv = lambda x: np.asarray(x) if isinstance(x, LazyByParticle) else x   # noqa: E731
        if "out" in kwargs:
            kwargs["out"] = tuple(conv(x) for x in kwargs["out"])
        return getattr(ufunc, method)(*(conv(x) for x in inputs), **kwargs)

    def __getitem__(self, item):
        return np.asarray(self)[item]

    def __len__(self):
        return self.shape[0]

    @property
    def size(self):
        return self.shape[0] * self.shape[1]

    @property
    def nbytes(self):
        return 8 * self.size

    def __getattr__(self, name):
        # only reached for names this class does not define: ndarray methods and attributes of the fetched array
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(np.asarray(self), name)

    def __repr__(self):
        state = "fetched" if self._cache is not None else ("stale" if self._stale else "on device")
        return f"<LazyByParticle shape={self.shape} float64, {state}>"


def regular_frame_window(analysis):
    """(first source frame, step) of the frames ``run()`` is about to visit when they form a regular progression with
    step >= 1, else ``None``.  MDAnalysis < 2.8 and the stand-in driver set ``start / stop / step``; MDAnalysis >= 2.8
    hands ``_setup_frames`` an explicit frame list (``start`` is then ``None``), so the list of the sliced trajectory is
    inspected instead."""
    n = int(getattr(analysis, "n_frames", 0) or 0)
    start, step = getattr(analysis, "start", None), getattr(analysis, "step", None)
    if start is not None and step is not None:
        return (int(start), int(step)) if step >= 1 else None
    sliced = getattr(analysis, "_sliced_trajectory", None)
    for name in ("frames", "_frames"):
        try:
            fr = getattr(sliced, name, None)
            if fr is None:
                continue
            fr = np.asarray(list(fr) if not isinstance(fr, np.ndarray) else fr)
            if fr.ndim != 1 or len(fr) != n or n == 0 or fr.dtype.kind not in "iu":
                continue
            if n == 1:
                return int(fr[0]), 1
            d = np.diff(fr)
            if d[0] >= 1 and np.all(d == d[0]):
                return int(fr[0]), int(d[0])
            return None
        except Exception:
            continue
    return None


class FrameStager:
    """Drives ta_stage_begin / slot / commit / bulk for one ``run()``."""

    def __init__(self, devices, n_frames: int, n_particles: int, dim_cols, n_fields: int,
                 masses, precision: str, pin_host: bool = True):
        # `devices`: device ids, or an already open Context to reuse (a second
        # run() on the same analysis object keeps its device buffers)
        if isinstance(devices, _lib.Context):
            self._devices, self._ctx = None, devices
        else:
            self._devices, self._ctx = devices, None
        self.T, self.N = int(n_frames), int(n_particles)
        self.dim_cols, self.n_fields = list(dim_cols), n_fields
        self.masses, self.precision = masses, precision
        self.pin_host = pin_host
        self._begun = False
        self._slab = None
        self._slab_frame0 = 0
        self._fill = 0
        self.bulk_done = False
        self.pinned = None            # bulk path: True if the reader's arrays are page-locked (DMA straight from them)

    @property
    def ctx(self) -> "_lib.Context":
        """The backend context, created on first use (so that data-presence
        errors surface before any device work, as in the reference); raises
        BackendError if the CUDA library or a GPU is missing."""
        if self._ctx is None:
            self._ctx = _lib.Context(self._devices)
        return self._ctx

    # -- whole-trajectory fast path ------------------------------------------
    def try_bulk(self, reader, atom_ix, window, need_positions: bool) -> bool:
        """Stream straight from an in-memory reader.  Conditions: 'fac' ordered float32 arrays, no on-the-fly
        transformations on the reader (they act on Timesteps, which this path never builds), a regular frame window
        (``window`` = (first, step) from :func:`regular_frame_window`) and a contiguous run of atoms; otherwise the
        per-frame path is used.  The arrays are page-locked once (``pin_host``; the registration is cached per array
        and dropped when the array is garbage-collected) so that the chunk copies are true asynchronous DMA; if the
        registration fails the copies go through the driver's staging buffers instead (slower, same results)."""
        if window is None:
            return False
        start, step = window
        if step < 1 or start < 0:
            return False
        if getattr(reader, "stored_order", None) != "fac":
            return False
        if len(tuple(getattr(reader, "transformations", ()) or ())) > 0:
            return False
        vel = getattr(reader, "velocity_array", None)
        if vel is None:
            return False
        fields = [vel]
        if need_positions:
            pos = reader.get_array() if hasattr(reader, "get_array") else getattr(reader, "coordinate_array", None)
            if pos is None:
                return False
            fields.append(pos)
        for a in fields:
            if not isinstance(a, np.ndarray) or a.dtype != np.float32 or a.ndim != 3 or not a.flags.c_contiguous:
                return False
            if a.shape[0] > (1 << 31) - 1 or start + (self.T - 1) * step >= a.shape[0]:
                return False
        ix = np.asarray(atom_ix)
        if len(ix) == 0 or not np.array_equal(ix, np.arange(ix[0], ix[0] + len(ix))):
            return False
        if self.pin_host:
            self.pinned = all([_lib.pin_array(a) for a in fields])
        else:
            self.pinned = all(_lib.is_pinned(a) for a in fields)
        self.ctx.stage_begin(self.T, self.N, self.dim_cols, np.float32, self.n_fields, self.masses, self.precision)
        self._begun = True
        self.ctx.stage_bulk(fields, atom_first=int(ix[0]), frame_first=int(start), frame_step=int(step),
                            nframes=self.T)
        self.bulk_done = True
        return True

    # -- per-frame path ----------------------------------------------------
    def add_frame(self, frame_index: int, velocities, positions=None, atom_ix=None):
        """Copy one frame into the current pinned slab.  ``velocities`` / ``positions``: the ``[N, 3]`` arrays of the
        atom group (what the reference copies, velocityautocorr.py:192-194), or -- with ``atom_ix`` -- the
        Timestep's full ``[n_atoms, 3]`` arrays, gathered straight into the slab (one pass instead of a temporary
        plus a copy)."""
        if not self._begun:
            dtype = np.float32 if velocities.dtype == np.float32 else np.float64
            self.ctx.stage_begin(self.T, self.N, self.dim_cols, dtype, self.n_fields, self.masses, self.precision)
            self._begun = True
        if self._slab is None:
            self._slab = self.ctx.stage_slot()
            self._slab_frame0 = frame_index
            self._fill = 0
        row = self._slab[self._fill]
        if atom_ix is None:
            row[0] = velocities
            if self.n_fields == 2:
                row[1] = positions
        else:
            _gather(velocities, atom_ix, row[0])
            if self.n_fields == 2:
                _gather(positions, atom_ix, row[1])
        self._fill += 1
        if self._fill == self._slab.shape[0]:
            self.flush()

    def flush(self):
        if self._slab is not None and self._fill > 0:
            self.ctx.stage_commit(self._slab_frame0, self._fill)
        self._slab = None
        self._fill = 0

    def finish(self):
        """Per-frame path: wait for the last slab.  Bulk path: nothing to wait for -- the compute call
        that follows is queued behind the staging chunk by chunk (H2D of chunk c+1 overlaps the
        correlation of chunk c) and synchronises at its end."""
        self.flush()
        if not self.bulk_done:
            self.ctx.stage_end()


def _gather(src, ix, out):
    """out[...] = src[ix] without a temporary: a slice copy for a contiguous run of atoms, else np.take."""
    if isinstance(ix, slice):
        out[...] = src[ix]
    else:
        np.take(src, ix, axis=0, out=out, mode="clip")


def gather_index(atom_ix, n_atoms_total=None):
    """Index for :func:`_gather`: a slice when the atoms are a contiguous run, else the index array."""
    ix = np.asarray(atom_ix)
    if len(ix) and np.array_equal(ix, np.arange(ix[0], ix[0] + len(ix))):
        return slice(int(ix[0]), int(ix[0]) + len(ix))
    return np.ascontiguousarray(ix, dtype=np.intp)
