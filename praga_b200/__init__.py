"""praga_b200 — B200-native gridding engine behind ARPA-SIMC/PRAGA's agrolib/interpolation API.

The product is praga_b200/libpraga_b200.so (hand-written sm_100a CUDA + the C ABI in include/praga_b200.h).
This package only holds the build script, the ctypes binding and the synthetic-scenario generator.
"""
from . import _abi as abi  # noqa: F401
from ._abi import Raster, Stations, default_proxy, default_settings  # noqa: F401
from .engine import Engine, PragaB200Error, empirical_variogram, fit_variogram, load_library  # noqa: F401

__all__ = ["abi", "Raster", "Stations", "default_proxy", "default_settings", "Engine", "PragaB200Error", "load_library",
           "empirical_variogram", "fit_variogram"]
