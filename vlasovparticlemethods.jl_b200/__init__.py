"""vpm_b200 — B200-native (sm_100a, fp64) particle hot path of VlasovMethods.jl behind a C ABI.

The directory is named after the reference repository (`vlasovparticlemethods.jl_b200`), which is not
an importable identifier; import it as `vpm_b200` through the loader module at the repository root.
"""
from . import _cabi
from ._cabi import LIB_PATH, VpmError, check
from .api import *  # noqa: F401,F403
from .fields import *  # noqa: F401,F403
from . import api, fields

__all__ = ([n for n in dir(api) if not n.startswith("_")] + [n for n in dir(fields) if not n.startswith("_")]
           + ["LIB_PATH", "VpmError", "check"])
