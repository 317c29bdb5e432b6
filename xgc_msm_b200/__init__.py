"""xgc_msm_b200 — B200-native (sm_100a) implementation of the weekly epidemic step of jrgant/xgc-msm.

The product is libxgc.so (CUDA kernels + C ABI, include/xgc/abi.h).  This package is its host-side mirror of the
reference's R interface (param_msm / control_msm / init_msm / netsim and the module(dat, at) -> dat contract).
Importing it loads libxgc.so and fails loudly when the library has not been built; there is no CPU fallback.
"""
from . import abi  # noqa: F401  (loads libxgc.so or raises ImportError)
from .abi import Handle, K, XgcError  # noqa: F401
from .params import control_msm, init_msm, param_msm  # noqa: F401

__all__ = ["abi", "Handle", "K", "XgcError", "param_msm", "control_msm", "init_msm"]
