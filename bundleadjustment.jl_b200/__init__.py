"""B200-native bundle-adjustment hot path: the BALNLPModel operator surface and the
Levenberg_Marquardt entry point of CelestineAngla/BundleAdjustment.jl, backed by hand-written
CUDA kernels for sm_100a behind the C ABI of include/bagpu.h (libbagpu.so).

The directory name carries a dot, so import it as ``bundleadjustment.jl_b200`` (the small
``bundleadjustment/`` shim at the repository root maps that dotted name onto this directory).
"""
from . import _lib
from ._lib import BAError
from .model import BALNLPModel, FeasibilityResidual, NLPModelMeta, Counters, name
from . import lm
from .lm import Levenberg_Marquardt, GenericExecutionStats, default_params, lm_step, last_solve_info
from . import synth
from .dist import init_comm

__all__ = ["BALNLPModel", "FeasibilityResidual", "NLPModelMeta", "Counters", "name", "Levenberg_Marquardt",
           "GenericExecutionStats", "default_params", "lm_step", "BAError", "synth", "init_comm"]
