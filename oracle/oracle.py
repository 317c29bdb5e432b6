"""TEST INFRASTRUCTURE — ctypes wrapper of the CPU oracle (oracle/libxgc_oracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module; the
product package xgc_msm_b200 never does.  PARITY UNPINNED: see oracle/xgc_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libxgc_oracle.so")


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("xgc_oracle.c", "xgc_oracle.h")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return LIB_PATH


class orc_inject(C.Structure):
    _fields_ = [("u", C.POINTER(C.c_double)), ("stride", C.c_size_t), ("uid_cap", C.c_int32),
                ("e_cap", C.c_int32 * 3), ("form_cap", C.c_int32)]


def _load():
    build()
    lib = C.CDLL(LIB_PATH)
    vp, ci, pd, pi, pu = C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint32)
    lib.orc_create.argtypes = [ci, C.POINTER(C.c_int), ci, C.c_uint64, ci]; lib.orc_create.restype = vp
    lib.orc_destroy.argtypes = [vp]; lib.orc_destroy.restype = None
    for f, a in {"orc_set_params": [vp, pd], "orc_set_control": [vp, pi], "orc_set_tables": [vp, pd],
                 "orc_set_population": [vp, ci, ci]}.items():
        getattr(lib, f).argtypes = a; getattr(lib, f).restype = None
    lib.orc_num_agents.argtypes = [vp]; lib.orc_num_agents.restype = ci
    lib.orc_set_attr.argtypes = [vp, C.c_char_p, pd, ci]; lib.orc_set_attr.restype = ci
    lib.orc_get_attr.argtypes = [vp, C.c_char_p, pd, ci]; lib.orc_get_attr.restype = ci
    lib.orc_set_edgelist.argtypes = [vp, ci, pi, ci, pi]; lib.orc_set_edgelist.restype = ci
    lib.orc_get_edgelist.argtypes = [vp, ci, pi, ci, pi]; lib.orc_get_edgelist.restype = ci
    lib.orc_get_actcounts.argtypes = [vp, ci, pi, ci]; lib.orc_get_actcounts.restype = ci
    lib.orc_get_actlist.argtypes = [vp, ci, pi, ci, C.POINTER(orc_inject)]; lib.orc_get_actlist.restype = ci
    lib.orc_step.argtypes = [vp, ci, C.c_uint32, C.POINTER(orc_inject)]; lib.orc_step.restype = None
    lib.orc_get_counters.argtypes = [vp, ci, pu]; lib.orc_get_counters.restype = None
    lib.orc_get_epi.argtypes = [vp, C.c_char_p, pd, ci, ci]; lib.orc_get_epi.restype = ci
    lib.orc_targets.argtypes = [vp, ci, ci, pd]; lib.orc_targets.restype = None
    lib.orc_status.argtypes = [vp]; lib.orc_status.restype = C.c_uint32
    lib.orc_philox4x32_10.argtypes = [pu, pu, pu]; lib.orc_philox4x32_10.restype = None
    lib.orc_exp.argtypes = [C.c_double]; lib.orc_exp.restype = C.c_double
    lib.orc_any_of_n.argtypes = [C.c_double, C.c_int]; lib.orc_any_of_n.restype = C.c_double
    lib.orc_bern.argtypes = [C.c_double, C.c_double]; lib.orc_bern.restype = ci
    lib.orc_pois.argtypes = [C.c_double, C.c_double, ci]; lib.orc_pois.restype = ci
    lib.orc_uniform.argtypes = [vp, ci, ci, C.c_uint32, C.c_uint32, ci]; lib.orc_uniform.restype = C.c_double
    lib.orc_constant.argtypes = [C.c_char_p, C.POINTER(C.c_int64)]; lib.orc_constant.restype = ci
    return lib


lib = _load()


def const(name: str) -> int:
    v = C.c_int64()
    if lib.orc_constant(("XGC_" + name).encode(), C.byref(v)):
        raise KeyError(name)
    return v.value


NPARAM, NCONTROL, NTABLE, NCOUNTER, NTARGET = (const(k) for k in ("NPARAM", "NCONTROL", "NTABLE", "NCOUNTER", "NTARGET"))


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Oracle:
    """One replicate of the CPU restatement."""

    def __init__(self, cap: int, e_cap, t_cap: int, seed: int = 1971, replicate_id: int = 0):
        ec = (C.c_int * 3)(*e_cap)
        self._d = lib.orc_create(cap, ec, t_cap, seed, replicate_id)
        self.cap, self.e_cap, self.t_cap = cap, list(e_cap), t_cap

    def close(self):
        if getattr(self, "_d", None):
            lib.orc_destroy(self._d)
            self._d = None

    __del__ = close

    def set_params(self, p):
        p = np.ascontiguousarray(p, dtype=np.float64); assert p.shape == (NPARAM,)
        lib.orc_set_params(self._d, _p(p, C.c_double))

    def set_control(self, c):
        c = np.ascontiguousarray(c, dtype=np.int32); assert c.shape == (NCONTROL,)
        lib.orc_set_control(self._d, _p(c, C.c_int32))

    def set_tables(self, t):
        t = np.ascontiguousarray(t, dtype=np.float64); assert t.shape == (NTABLE,)
        lib.orc_set_tables(self._d, _p(t, C.c_double))

    def set_population(self, n: int, at: int = 1):
        lib.orc_set_population(self._d, n, at)

    def num_agents(self) -> int:
        return lib.orc_num_agents(self._d)

    def set_attr(self, name: str, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        if lib.orc_set_attr(self._d, name.encode(), _p(v, C.c_double), v.size):
            raise KeyError(f"oracle: bad attribute '{name}' or length {v.size}")

    def get_attr(self, name: str) -> np.ndarray:
        v = np.empty(self.cap)
        n = lib.orc_get_attr(self._d, name.encode(), _p(v, C.c_double), self.cap)
        if n < 0:
            raise KeyError(name)
        return v[:n].copy()

    def set_edgelist(self, layer: int, el, start=None):
        el = np.asarray(el, dtype=np.int32).reshape(-1, 2)
        cm = np.ascontiguousarray(el.T)
        st = None if start is None else np.ascontiguousarray(start, dtype=np.int32)
        if lib.orc_set_edgelist(self._d, layer, _p(cm, C.c_int32), el.shape[0], None if st is None else _p(st, C.c_int32)):
            raise ValueError("oracle: too many edges")

    def get_edgelist(self, layer: int):
        cap = self.e_cap[layer]
        flat = np.empty(2 * cap, dtype=np.int32); st = np.empty(cap, dtype=np.int32)
        ne = lib.orc_get_edgelist(self._d, layer, _p(flat, C.c_int32), cap, _p(st, C.c_int32))
        return np.stack([flat[:ne], flat[ne:2 * ne]], axis=1).copy(), st[:ne].copy()

    def get_actcounts(self, layer: int) -> np.ndarray:
        cap = self.e_cap[layer]
        flat = np.empty(4 * cap, dtype=np.int32)
        ne = lib.orc_get_actcounts(self._d, layer, _p(flat, C.c_int32), cap)
        return flat[: 4 * ne].reshape(4, ne).T.copy()

    def get_actlist(self, at: int, cap: int = 1 << 16, inj: Optional[orc_inject] = None) -> np.ndarray:
        al = np.empty(6 * cap, dtype=np.int32)
        n = lib.orc_get_actlist(self._d, at, _p(al, C.c_int32), cap, None if inj is None else C.byref(inj))
        if n < 0:
            raise ValueError("act list buffer too small")
        return al.reshape(6, cap)[:, :n].T.copy()

    def step(self, at: int, mask: int, inj: Optional[orc_inject] = None):
        lib.orc_step(self._d, at, mask, None if inj is None else C.byref(inj))

    def counters(self, at: int) -> np.ndarray:
        out = np.empty(NCOUNTER, dtype=np.uint32)
        lib.orc_get_counters(self._d, at, _p(out, C.c_uint32))
        return out

    def epi(self, var: str, at0: int, at1: int) -> np.ndarray:
        out = np.empty(at1 - at0 + 1)
        if lib.orc_get_epi(self._d, var.encode(), _p(out, C.c_double), at0, at1):
            raise KeyError(var)
        return out

    def targets(self, at_last: int, tailn: int) -> np.ndarray:
        out = np.empty(NTARGET)
        lib.orc_targets(self._d, at_last, tailn, _p(out, C.c_double))
        return out

    def status(self) -> int:
        return lib.orc_status(self._d)


def make_inject(u: Optional[np.ndarray], stride: int, uid_cap: int, e_cap, form_cap: int) -> orc_inject:
    j = orc_inject()
    j.stride, j.uid_cap, j.form_cap = stride, uid_cap, form_cap
    j.e_cap[:] = list(e_cap)
    if u is not None:
        assert u.dtype == np.float64 and u.flags.c_contiguous and u.size == stride
        j.u = _p(u, C.c_double)
    return j
