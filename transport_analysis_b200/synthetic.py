"""Synthetic in-memory universes (harness for tests and bench, not hot path).

`make_universe` follows the reference's recipe for MemoryReader-backed dummy
universes (transport_analysis/tests/utils.py:8-77): float ``[frames, atoms, 3]``
arrays handed to a MemoryReader.  With MDAnalysis installed a real
``mda.Universe`` is built; otherwise the stand-in from ``_compat`` is used.
"""
from __future__ import annotations

import numpy as np

from . import _compat


def make_universe(positions=None, velocities=None, masses=None, dimensions=None, dt=1.0):
    if _compat.HAVE_MDANALYSIS:  # pragma: no cover
        import MDAnalysis as mda
        from MDAnalysis.coordinates.memory import MemoryReader

        ref = positions if positions is not None else velocities
        n_frames, n_atoms = ref.shape[0], ref.shape[1]
        u = mda.Universe.empty(n_atoms, trajectory=True, velocities=velocities is not None)
        pos = positions if positions is not None else np.zeros_like(velocities)
        dims = None
        if dimensions is not None:
            dims = np.asarray(dimensions, dtype=np.float32)
            if dims.ndim == 1:
                dims = np.tile(dims, (n_frames, 1))
        u.trajectory = MemoryReader(np.asarray(pos, dtype=np.float32), velocities=velocities,
                                    dimensions=dims, dt=dt)
        if masses is not None:
            u.add_TopologyAttr("masses", np.asarray(masses, dtype=np.float64))
        return u
    return _compat.Universe(positions=positions, velocities=velocities, masses=masses,
                            dimensions=dimensions, dt=dt)


def random_trajectory(n_frames, n_atoms, seed=0, with_positions=False, rho=0.0, box=20.0, dt=1.0):
    """Seeded synthetic trajectory: velocities ~ N(0, 1),
    optionally AR(1)-correlated in time with coefficient ``rho``; positions =
    U(0, box) + cumulative sum of v*dt (unwrapped).  Values are float32 (what
    an MDAnalysis Timestep carries).  Returns (velocities, positions|None)."""
    rng = np.random.default_rng(seed)
    vel = rng.standard_normal((n_frames, n_atoms, 3))
    if rho:
        for t in range(1, n_frames):
            vel[t] = rho * vel[t - 1] + np.sqrt(1 - rho * rho) * vel[t]
    pos = None
    if with_positions:
        pos = rng.uniform(0, box, (1, n_atoms, 3)) + np.cumsum(vel * dt, axis=0)
        pos = pos.astype(np.float32)
    return vel.astype(np.float32), pos
