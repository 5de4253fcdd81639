"""vbmf_b200: B200-native VB matrix-factorisation update loop behind VBMatrixFactorization.jl's API.

The package directory is named `vbmatrixfactorization.jl_b200` (not importable by that dotted name); load it as module
`vbmf_b200` through `vbmf_b200_loader.load()` at the repository root.
"""
from . import _lib
from ._lib import VBMFError, LIB_PATH
from .api import *  # noqa: F401,F403
from .api import Solver
from . import mil  # noqa: F401  (callers of the path in the reference's MIL example)
