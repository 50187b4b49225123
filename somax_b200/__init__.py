"""somax_b200 - B200-native time-stepping hot path of jejjohnson/somax.

Hand-written CUDA (sm_100a) behind a C ABI (``include/somax_b200.h``); this package is the
host-side mirror of the reference's equinox model API for that path.  There is no CPU
fallback: compute calls raise if ``libsomax_b200.so`` is missing or no B200 is visible.
"""
from . import _lib  # noqa: F401
from .core import (  # noqa: F401
    Diagnostics, ModalTransform, ODETerm, Params, PhysConsts, SaveAt, Solution, SomaxModel, State,
    StratificationProfile, Grid,
)
from .models.qg import (  # noqa: F401
    BaroclinicQG, BaroclinicQGDiagnostics, BaroclinicQGParams, BaroclinicQGPhysConsts,
    BaroclinicQGState, BarotropicQG, BarotropicQGDiagnostics, BarotropicQGParams,
    BarotropicQGPhysConsts, BarotropicQGState,
)
from .models.swm import (  # noqa: F401
    MultilayerShallowWater2D, MultilayerSW2DDiagnostics, MultilayerSW2DParams,
    MultilayerSW2DPhysConsts, MultilayerSW2DState, NonlinearShallowWater2D,
    NonlinearSW2DDiagnostics, NonlinearSW2DParams, NonlinearSW2DPhysConsts, NonlinearSW2DState,
)
from .models.reparam import ReparameterizedQG, ReparamQGDiagnostics  # noqa: F401
from . import gfd_testcases  # noqa: F401

__version__ = "0.1.0"
