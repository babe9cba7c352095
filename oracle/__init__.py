"""CPU oracle for the time-correlation hot path (TEST INFRASTRUCTURE ONLY).

This package is a numpy restatement of the reference's algorithm for the
hot path named by BASELINE.json (`VelocityAutocorr._conclude_fft`,
`VelocityAutocorr._conclude_simple`, `ViscosityHelfand._conclude`, and the
un-vendored `tidynamics.acf` they call).  It exists to CHECK the CUDA path:

* only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
  ``cpu_baseline`` / ``--impl reference`` legs may import it;
* nothing under ``transport_analysis_b200/`` imports it, and the product
  path raises if the CUDA library is missing -- there is no CPU fallback.

Parity status: PINNED.  The restatement is checked in ``tests/test_oracle.py``
against (a) the reference's closed-form known answers
(`characteristic_poly`, `characteristic_poly_helfand`), (b) the printed
outputs of the reference's tutorial notebooks, and (c) golden vectors made by
executing the reference's own ``_conclude*`` code in this container
(``tests/golden/make_golden.py``).
"""
from .reference_numpy import (  # noqa: F401
    BOLTZMANN_KJ_PER_MOL_K,
    parse_dim_type,
    tidynamics_acf,
    vacf_fft,
    vacf_windowed,
    helfand_msd,
    polyfit_viscosity,
)
from .known_answers import (  # noqa: F401
    characteristic_poly,
    characteristic_poly_helfand,
    NOTEBOOK_VACF_T10_XYZ,
    NOTEBOOK_HELFAND_T10_SUMDIMS,
)
