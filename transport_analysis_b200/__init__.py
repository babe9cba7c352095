"""transport_analysis_b200 -- B200-native backend for the time-correlation hot
path of MDAnalysis/transport-analysis (VACF, FFT and windowed; Helfand MSD).

The public classes mirror the reference package:

>>> from transport_analysis_b200.velocityautocorr import VelocityAutocorr
>>> from transport_analysis_b200.viscosity import ViscosityHelfand
"""
from .velocityautocorr import VelocityAutocorr  # noqa: F401
from .viscosity import ViscosityHelfand  # noqa: F401
from ._lib import BackendError  # noqa: F401

__version__ = "0.1.0"
