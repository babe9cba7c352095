"""Velocity autocorrelation function on B200 GPUs.

Drop-in for ``transport_analysis.velocityautocorr.VelocityAutocorr``
(reference: transport_analysis/velocityautocorr.py:72-422): same constructor
arguments, ``dim_type``, ``fft``, ``run(start, stop, step)``,
``results.timeseries`` / ``results.vacf_by_particle`` and the Green-Kubo
helpers.  ``_single_frame`` stages frames into pinned slabs that are streamed
to HBM while the trajectory loop runs; ``_conclude`` is one call into
``libta_b200.so`` (FFT route: kernel K1; windowed route: kernel K2).
There is no CPU fallback.

Extra keyword arguments (all default to the reference's behaviour):

``precision``  ``"fp64"`` (default, matches the reference to 1e-10) or
               ``"fp32"`` (stated tolerance 1e-5).
``devices``    CUDA device ids the particles are sharded over (default ``[0]``
               or ``$TA_B200_DEVICES``).
``max_eager_bytes``  per-particle results larger than this stay on the GPUs
               behind a lazy handle that behaves like the array and is fetched on first use (default 64 MiB).
``staging``    ``"auto"`` (default): in-memory readers (MemoryReader) are streamed to the GPU as whole arrays,
               every other reader frame by frame through pinned slabs.  ``"per_frame"``: always frame by frame
               (the path TRR / XTC / NetCDF readers take).
``postprocess``  ``"host"`` (default): the Green-Kubo helpers are the reference's scipy calls on the host timeseries.
               ``"device"``: the trapezoid integral and the running integral are computed on the GPU from the
               timeseries that is still there (kernel K7); Simpson's rule stays on the host.
``pin_host``   ``True`` (default): the arrays of an in-memory reader are page-locked on first use so that their
               copies are asynchronous DMA (a one-off cost of ~0.2 s per GB); ``False`` leaves them pageable.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from ._compat import AnalysisBase, NoDataError, UpdatingAtomGroup
from ._staging import FrameStager, LazyByParticle, gather_index, regular_frame_window, resolve_devices

_DIM_KEYS = {
    "x": [0],
    "y": [1],
    "z": [2],
    "xy": [0, 1],
    "xz": [0, 2],
    "yz": [1, 2],
    "xyz": [0, 1, 2],
}


def parse_dim_type(dim_str):
    """``dim_type`` -> (columns, dimensionality); same error text as the
    reference (velocityautocorr.py:155-176)."""
    try:
        cols = _DIM_KEYS[dim_str]
    except KeyError:
        raise ValueError(
            "invalid dim_type: {} specified, please specify one of xyz, "
            "xy, xz, yz, x, y, z".format(dim_str)
        )
    return cols, len(cols)


class VelocityAutocorr(AnalysisBase):
    """Per-particle VACF, averaged over the atom group.

    Parameters
    ----------
    atomgroup : AtomGroup (``UpdatingAtomGroup`` is rejected)
    dim_type : {'xyz', 'xy', 'yz', 'xz', 'x', 'y', 'z'}
    fft : bool -- ``True``: FFT route (replaces the tidynamics.acf loop);
        ``False``: the windowed lag sums.

    Attributes set by :meth:`run`: ``results.timeseries`` (``[n_frames]``),
    ``results.vacf_by_particle`` (``[n_frames, n_particles]``), ``times``,
    ``n_frames``, ``n_particles``, ``dim_fac``.
    """

    def __init__(self, atomgroup, dim_type="xyz", fft=True, precision="fp64", devices=None,
                 max_eager_bytes=1 << 26, staging="auto", pin_host=True, postprocess="host", **kwargs):
        super().__init__(atomgroup.universe.trajectory, **kwargs)

        if isinstance(atomgroup, UpdatingAtomGroup):
            raise TypeError("UpdatingAtomGroups are not valid for VACF computation")
        if postprocess not in ("host", "device"):
            raise ValueError("postprocess must be 'host' or 'device'")
        self._postprocess = postprocess

        self.dim_type = dim_type.lower()
        self._dim, self.dim_fac = parse_dim_type(self.dim_type)
        self.fft = fft
        if precision not in ("fp64", "fp32"):
            raise ValueError("precision must be 'fp64' or 'fp32'")
        self.precision = precision
        if staging not in ("auto", "per_frame"):
            raise ValueError("staging must be 'auto' or 'per_frame'")
        self._staging, self._pin_host = staging, bool(pin_host)
        self._devices = resolve_devices(devices)
        self._max_eager_bytes = int(max_eager_bytes)

        self.atomgroup = atomgroup
        self.n_particles = len(self.atomgroup)
        self._run_called = False
        self._ctx = None

    _parse_dim_type = staticmethod(parse_dim_type)

    # -- AnalysisBase hooks --------------------------------------------------
    def _prepare(self):
        if self.n_frames < 1 or self.n_particles < 1:
            raise ValueError("VACF needs at least one frame and one particle")
        prev = self.results.get("vacf_by_particle") if hasattr(self.results, "get") else None
        if isinstance(prev, LazyByParticle):
            prev.invalidate()             # the device buffers are about to be reused
        self._stager = FrameStager(self._ctx or self._devices, self.n_frames, self.n_particles, self._dim, 1, None,
                                   self.precision, self._pin_host)
        self._gather_ix = gather_index(self.atomgroup.ix)
        if self._staging == "auto":
            self._stager.try_bulk(self._trajectory, self.atomgroup.ix, regular_frame_window(self), False)

    def _single_frame(self):
        ts = self._ts
        if not ts.has_velocities:
            raise NoDataError("VACF computation requires velocities in the trajectory")
        if self._stager.bulk_done:
            return
        # atomgroup.velocities is ts.velocities[atomgroup.ix] (a temporary); gather straight into the pinned slab instead
        self._stager.add_frame(self._frame_index, ts.velocities, atom_ix=self._gather_ix)

    def _conclude(self):
        self._stager.finish()
        self._ctx = self._stager.ctx
        if self.fft:
            self.results.timeseries = self._ctx.vacf_fft()
        else:
            self.results.timeseries = self._ctx.vacf_windowed()
        nbytes = 8 * self.n_frames * self.n_particles
        if nbytes <= self._max_eager_bytes:
            self.results.vacf_by_particle = self._ctx.fetch_by_particle()
        else:
            self.results.vacf_by_particle = LazyByParticle(self._ctx, self.n_frames, self.n_particles)
        self._run_called = True

    # -- Green-Kubo helpers (host-side, O(T); reference :240-422) -----------------
    def _window(self, start, stop, step, what):
        if not self._run_called:
            raise RuntimeError(f"Analysis must be run prior to {what}")
        stop = self.n_frames if stop == 0 else stop
        return slice(start, stop, step)

    def self_diffusivity_gk(self, start=0, stop=0, step=1):
        """Trapezoid-rule Green-Kubo self-diffusivity (reference :287-322)."""
        from scipy import integrate

        w = self._window(start, stop, step, "computing self-diffusivity")
        if self._postprocess == "device":
            lo, hi, st = w.indices(self.n_frames)
            return self._ctx.green_kubo(self.times, lo, max(lo, hi), st)[0] / self.dim_fac
        return integrate.trapezoid(self.results.timeseries[w], self.times[w]) / self.dim_fac

    def self_diffusivity_gk_odd(self, start=0, stop=0, step=1):
        """Simpson-rule Green-Kubo self-diffusivity (reference :324-360)."""
        from scipy import integrate

        w = self._window(start, stop, step, "computing self-diffusivity")
        return integrate.simpson(y=self.results.timeseries[w], x=self.times[w]) / self.dim_fac

    def running_integral(self, start=0, stop=0, step=1, initial=0):
        """(times, cumulative trapezoid of the VACF / dim_fac): the data behind
        :meth:`plot_running_integral` (reference :407-414)."""
        from scipy import integrate

        w = self._window(start, stop, step, "plotting")
        if self._postprocess == "device":
            lo, hi, st = w.indices(self.n_frames)
            vals = self._ctx.green_kubo(self.times, lo, max(lo, hi), st, initial=initial, running=True)[2]
        else:
            vals = integrate.cumulative_trapezoid(self.results.timeseries[w], self.times[w], initial=initial)
        return self.times[w], vals / self.dim_fac

    def plot_vacf(self, start=0, stop=0, step=1, xlabel="Time (ps)",
                  ylabel="Velocity Autocorrelation Function (Å^2 / ps^2)"):
        """Matplotlib line of the VACF (reference :240-285)."""
        w = self._window(start, stop, step, "plotting")
        import matplotlib.pyplot as plt

        _, ax = plt.subplots()
        ax.set_xlabel(xlabel)
        ax.set_ylabel(ylabel)
        return ax.plot(self.times[w], self.results.timeseries[w])

    def plot_running_integral(self, start=0, stop=0, step=1, initial=0, xlabel="Time (ps)",
                              ylabel="Running Integral of the VACF (Å^2 / ps)"):
        """Matplotlib line of the running integral (reference :362-422)."""
        times, vals = self.running_integral(start, stop, step, initial)
        import matplotlib.pyplot as plt

        _, ax = plt.subplots()
        ax.set_xlabel(xlabel)
        ax.set_ylabel(ylabel)
        return ax.plot(times, vals)
