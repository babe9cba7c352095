"""Generate golden vectors by EXECUTING THE REFERENCE'S OWN CODE.

Runs only in the build container (needs /root/reference); the GPU box and the
test-suite read the committed ``tests/golden/*.npz`` instead.

MDAnalysis, tidynamics and matplotlib are not installable here, so the
reference modules are imported with stand-ins injected into ``sys.modules``:

* ``MDAnalysis.analysis.base.AnalysisBase`` & friends -> the protocol
  stand-in from ``transport_analysis_b200._compat`` (driver only);
* ``tidynamics.acf`` -> ``oracle.tidynamics_acf`` (the restated third-party
  algorithm; vectors that went through it are tagged ``*_fft``);
* ``matplotlib`` -> empty stubs.

Everything else -- ``_prepare``, ``_single_frame``, ``_conclude_simple``,
``_conclude_fft``'s per-particle loop and mean, ``ViscosityHelfand._conclude``
-- is the reference's unmodified source executed by numpy.  The windowed VACF
and Helfand vectors therefore pin the oracle against the real reference.

Usage:  python tests/golden/make_golden.py
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REFERENCE = "/root/reference"


def _inject_stubs():
    from transport_analysis_b200 import _compat
    import oracle

    assert not _compat.HAVE_MDANALYSIS

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("MDAnalysis")
    mod("MDAnalysis.analysis")
    mod("MDAnalysis.analysis.base", AnalysisBase=_compat.AnalysisBase, Results=_compat.Results)
    mod("MDAnalysis.core")
    mod("MDAnalysis.core.groups", UpdatingAtomGroup=_compat.UpdatingAtomGroup)
    mod("MDAnalysis.exceptions", NoDataError=_compat.NoDataError)
    mod("MDAnalysis.units", constants=_compat.constants)
    mod("tidynamics", acf=oracle.tidynamics_acf)
    mod("matplotlib")
    mod("matplotlib.pyplot")
    # import the two modules without running the package __init__ (versioneer)
    pkg = types.ModuleType("transport_analysis")
    pkg.__path__ = [os.path.join(REFERENCE, "transport_analysis")]
    sys.modules["transport_analysis"] = pkg


def main():
    _inject_stubs()
    from transport_analysis.velocityautocorr import VelocityAutocorr as RefVACF  # noqa: E402
    from transport_analysis.viscosity import ViscosityHelfand as RefVH  # noqa: E402
    from transport_analysis_b200.synthetic import make_universe, random_trajectory

    out = {}
    # --- random trajectory, small enough to commit (float32 inputs are stored)
    T, N = 160, 6
    vel, pos = random_trajectory(T, N, seed=7, with_positions=True, rho=0.8)
    masses = np.array([1.008, 12.011, 15.999, 1.008, 12.011, 15.999])
    box = np.tile(np.array([20, 20, 20, 90, 90, 90], dtype=np.float32), (T, 1))
    box[:, 0] *= 1 + 0.01 * np.sin(np.arange(T))          # +-1 % volume jitter
    u = make_universe(pos, vel, masses=masses, dimensions=box)
    out["rand_vel"], out["rand_pos"], out["rand_masses"], out["rand_box"] = vel, pos, masses, box
    for dim in ("xyz", "xy", "xz", "yz", "x", "y", "z"):
        r = RefVACF(u.atoms, dim_type=dim, fft=False).run()
        out[f"rand_vacf_windowed_{dim}_ts"] = r.results.timeseries
        out[f"rand_vacf_windowed_{dim}_bp"] = r.results.vacf_by_particle
        r = RefVACF(u.atoms, dim_type=dim, fft=True).run()
        out[f"rand_vacf_fft_{dim}_ts"] = r.results.timeseries
        out[f"rand_vacf_fft_{dim}_bp"] = r.results.vacf_by_particle
        r = RefVH(u.atoms, temp_avg=310.0, dim_type=dim).run()
        out[f"rand_helfand_{dim}_ts"] = r.results.timeseries
        out[f"rand_helfand_{dim}_bp"] = r.results.visc_by_particle
    # sliced run + atom subset + linear fit
    r = RefVACF(u.atoms[1:5], fft=False).run(start=3, stop=150, step=4)
    out["rand_vacf_windowed_sliced_ts"] = r.results.timeseries
    out["rand_vacf_windowed_sliced_times"] = r.times
    r = RefVACF(u.atoms[1:5], fft=True).run(start=3, stop=150, step=4)
    out["rand_vacf_fft_sliced_ts"] = r.results.timeseries
    r = RefVH(u.atoms[1:5], linear_fit_window=(5, 30)).run(start=3, stop=150, step=4)
    out["rand_helfand_sliced_ts"] = r.results.timeseries
    out["rand_helfand_sliced_viscosity"] = np.float64(r.results.viscosity)
    # GK helpers on the full windowed xyz run
    r = RefVACF(u.atoms, fft=False).run()
    out["rand_gk"] = np.float64(r.self_diffusivity_gk())
    out["rand_gk_odd"] = np.float64(r.self_diffusivity_gk_odd(start=0, stop=159))
    out["rand_gk_sliced"] = np.float64(r.self_diffusivity_gk(start=2, stop=100, step=3))

    # --- the reference's step trajectory (tests/test_velocityautocorr.py:46-57,
    #     tests/test_viscosity.py:59-86), short version; full length is checked
    #     against the closed forms in tests/test_oracle.py
    NS = 301
    t = np.arange(NS, dtype=np.float64)
    v = np.repeat(t[:, None, None], 3, axis=2)
    x = np.repeat((t * t / 2)[:, None, None], 3, axis=2)
    us = make_universe(x, v, masses=[16.0], dimensions=[2, 2, 2, 90, 90, 90])
    for dim in ("xyz", "xy", "x"):
        out[f"step_vacf_windowed_{dim}_ts"] = RefVACF(us.atoms, dim_type=dim, fft=False).run().results.timeseries
        out[f"step_vacf_fft_{dim}_ts"] = RefVACF(us.atoms, dim_type=dim, fft=True).run().results.timeseries
        out[f"step_helfand_{dim}_ts"] = RefVH(us.atoms, dim_type=dim).run().results.timeseries
    out["step_helfand_sliced_ts"] = RefVH(us.atoms).run(start=10, stop=300, step=10).results.timeseries

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_run.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
