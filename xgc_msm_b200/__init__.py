"""xgc_msm_b200 — B200-native (sm_100a) implementation of the weekly epidemic step of jrgant/xgc-msm.

The product is libxgc.so (CUDA kernels + C ABI, include/xgc/abi.h).  This package is its host-side mirror of the
reference's R interface (param_msm / control_msm / init_msm / netsim and the module(dat, at) -> dat contract).
Anything that computes goes through `abi` (ctypes binding of libxgc.so), which fails loudly when the library has not
been built; there is no CPU fallback.  The binding is loaded on first use (`xgc_msm_b200.abi`, `Handle`, `K`), not at
package import, so that the pure-Python workload description (params / synth / workload) can be imported by the CPU
arms of bench.py without mapping the CUDA library into those processes.
"""
from .params import control_msm, init_msm, param_msm  # noqa: F401

__all__ = ["abi", "Handle", "K", "XgcError", "param_msm", "control_msm", "init_msm"]


def __getattr__(name):
    if name in ("abi", "Handle", "K", "XgcError"):
        import importlib
        abi = importlib.import_module(".abi", __name__)          # loads libxgc.so or raises ImportError
        return abi if name == "abi" else getattr(abi, name)
    raise AttributeError(name)
