"""Import alias: the product package lives in ``matrixproductbp.jl_b200/`` (a directory name Python
cannot import directly because of the dot); ``import mpbp_b200`` exposes it."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "..", "matrixproductbp.jl_b200"))
from .api import *  # noqa: E402,F401,F403
from . import api as _api  # noqa: E402

__all__ = _api.__all__
