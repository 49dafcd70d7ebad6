"""klara_b200: B200-native (sm_100a) implementation of the MCMC inner loop of JuliaStats/Klara.jl
(HMC leapfrog / MALA drift proposal / MH accept-reject, batched over independent chains), behind
Klara's own user surface.  See DESIGN.md and INTEGRATION.md.

The directory is named `klara.jl_b200`; because of the dot it is imported through the
`klara_b200` loader module at the repo root (`import klara_b200`).
"""
from . import _lib, distributed, gibbs, iostream
from .gibbs import BasicGibbsJob, Transformation
from .iostream import BasicContParamIOStream
from ._lib import KlaraError
from .api import *  # noqa: F401,F403
from .api import __all__ as _api_all
from .targets import BayesLogit, DenseGaussian, IsoGaussian, Rosenbrock, ShiftedIsoGaussian, Target

__all__ = list(_api_all) + ["IsoGaussian", "ShiftedIsoGaussian", "Rosenbrock", "DenseGaussian", "BayesLogit", "Target",
                            "KlaraError", "BasicContParamIOStream", "BasicGibbsJob", "Transformation"]
