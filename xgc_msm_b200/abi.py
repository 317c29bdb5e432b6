"""ctypes binding of the C ABI in include/xgc/abi.h (libxgc.so).

This is the same boundary an R maintainer binds with `.Call`; Python is used here because the
reference's own host language (R) is not installed in this image.  There is no CPU fallback: if libxgc.so is missing
the import fails, and if no GPU is present `Handle()` raises with the library's error message.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libxgc.so")


class XgcError(RuntimeError):
    pass


class xgc_config(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("n_replicates", C.c_int32),
                ("replicate_offset", C.c_int32), ("n_cap", C.c_int32), ("e_cap", C.c_int32 * 3),
                ("t_cap", C.c_int32), ("compact_every", C.c_int32), ("seed", C.c_uint64)]


class xgc_inject(C.Structure):
    _fields_ = [("u", C.POINTER(C.c_double)), ("stride", C.c_size_t), ("uid_cap", C.c_int32),
                ("e_cap", C.c_int32 * 3), ("form_cap", C.c_int32)]


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: build it with ./build.sh or __graft_entry__.build(); "
                          "xgc_msm_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp, ci, cd, cu = C.c_void_p, C.c_int, C.c_double, C.c_uint32
    pd, pi, pu = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint32)
    sig = {
        "xgc_create": [C.POINTER(xgc_config), C.POINTER(vp)], "xgc_destroy": [vp], "xgc_abi_version": [],
        "xgc_set_params": [vp, ci, pd], "xgc_get_params": [vp, ci, pd], "xgc_set_params_all": [vp, pd],
        "xgc_set_control": [vp, ci, pi], "xgc_set_tables": [vp, pd],
        "xgc_set_mode": [vp, ci], "xgc_get_mode": [vp, pi],
        "xgc_get_tables": [vp, pd], "xgc_get_params_all": [vp, pd], "xgc_get_control": [vp, ci, pi],
        "xgc_set_population": [vp, ci, ci, ci], "xgc_num_agents": [vp, ci, pi],
        "xgc_upload_attr": [vp, ci, C.c_char_p, vp, C.c_size_t, ci],
        "xgc_download_attr": [vp, ci, C.c_char_p, vp, C.c_size_t, ci, C.POINTER(C.c_size_t)],
        "xgc_attr_names": [C.POINTER(C.c_char_p), ci], "xgc_epi_names": [C.POINTER(C.c_char_p), ci],
        "xgc_set_edgelist": [vp, ci, ci, pi, ci, pi], "xgc_get_edgelist": [vp, ci, ci, pi, ci, pi, pi],
        "xgc_get_actcounts": [vp, ci, ci, pi, ci, pi],
        "xgc_get_actlist": [vp, ci, ci, pi, ci, pi, C.POINTER(xgc_inject)],
        "xgc_step": [vp, ci, cu, C.POINTER(xgc_inject)], "xgc_step_span": [vp, ci, cu, cu], "xgc_set_replicate_ids": [vp, pi], "xgc_replicate_cost": [vp, C.POINTER(C.c_uint64), ci], "xgc_week_hostnet": [vp, ci, cu, cu, vp, ci, vp, ci, ci, vp, vp, ci, vp, vp], "xgc_run": [vp, ci, ci], "xgc_sync": [vp],
        "xgc_compact": [vp], "xgc_status": [vp, ci, pu],
        "xgc_get_counters": [vp, ci, ci, ci, pu], "xgc_get_epi": [vp, ci, C.c_char_p, pd, ci, ci],
        "xgc_get_epi_table": [vp, ci, C.POINTER(C.c_char_p), ci, pd, ci, ci],
        "xgc_targets": [vp, ci, ci, pd],
        "xgc_snapshot_size": [vp, C.POINTER(C.c_size_t)], "xgc_snapshot": [vp, vp, C.c_size_t],
        "xgc_restore": [vp, vp, C.c_size_t], "xgc_fork": [vp, ci, vp, ci],
        "xgc_set_edgelists_all": [vp, ci, vp, vp, vp, ci], "xgc_update_edgelists_all": [vp, ci, vp, ci, ci], "xgc_get_counters_all": [vp, ci, vp],
        "xgc_kernel_launches": [vp, C.POINTER(C.c_uint64)], "xgc_profile": [vp, ci],
        "xgc_profile_read": [vp, C.POINTER(C.c_char_p), pd, C.POINTER(C.c_uint64), ci], "xgc_num_agents_all": [vp, vp], "xgc_stream_ptr": [vp, C.POINTER(vp)],
        "xgc_debug_exp": [pd, pd, ci, ci], "xgc_debug_philox": [cu, cu, pu, pu, ci, ci],
        "xgc_constant": [C.c_char_p, C.POINTER(C.c_int64)], "xgc_param_index": [C.c_char_p, pi, pi],
        "xgc_inject_stream_width": [ci],
    }
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = ci
    lib.xgc_last_error.argtypes = [vp]
    lib.xgc_last_error.restype = C.c_char_p
    return lib


lib = _load()


class _Constants:
    """XGC_* constants of abi.h, resolved through xgc_constant() (K.NPARAM, K.M_ALL, K.S_ACTS ...)."""

    def __getattr__(self, name: str) -> int:
        v = C.c_int64()
        if lib.xgc_constant(("XGC_" + name).encode(), C.byref(v)) != 0:
            raise AttributeError(name)
        setattr(self, name, v.value)
        return v.value


K = _Constants()


def names(which: str) -> list[str]:
    fn = {"attr": lib.xgc_attr_names, "epi": lib.xgc_epi_names}[which]
    n = fn(None, 0)
    buf = (C.c_char_p * n)()
    fn(buf, n)
    return [b.decode() for b in buf]


def _ptr(a: np.ndarray, typ):
    return a.ctypes.data_as(C.POINTER(typ))


def make_inject(u: Optional[np.ndarray], uid_cap: int, e_cap, form_cap: int) -> xgc_inject:
    """Describe an injected-draw table u[n_replicates, stride] (float64, C order); stride must equal inject_size()."""
    j = xgc_inject()
    j.uid_cap = uid_cap
    j.e_cap[:] = list(e_cap)
    j.form_cap = form_cap
    j.stride = inject_size(uid_cap, e_cap, form_cap)
    if u is not None:
        assert u.dtype == np.float64 and u.flags.c_contiguous and u.shape[-1] == j.stride
        j.u = _ptr(u, C.c_double)
    return j


def stream_width(s: int) -> int:
    """Draws per entity of stream s (the C header's xgc_stream_width, through the library so the two cannot drift)."""
    w = lib.xgc_inject_stream_width(int(s))
    if w < 0:
        raise ValueError(f"no such stream {s}")
    return w


def stream_entities(s: int, uid_cap: int, e_cap, form_cap: int) -> int:
    if s <= K.S_ARRIVE_ATTR:
        return uid_cap
    if s < K.S_FORM:
        return 1
    if s < K.S_EDGE0:
        return form_cap
    return e_cap[(s - K.S_EDGE0) // K.EDGE_STREAMS]


def inject_size(uid_cap: int, e_cap, form_cap: int) -> int:
    return sum(stream_entities(s, uid_cap, e_cap, form_cap) * stream_width(s) for s in range(K.NSTREAM))


class Handle:
    """One GPU's shard of replicates (opaque xgc_handle*)."""

    def __init__(self, n_replicates: int, n_cap: int, e_cap, t_cap: int, seed: int = 1971, device: int = 0,
                 replicate_offset: int = 0, compact_every: int = 0):
        cfg = xgc_config()
        cfg.abi_version = K.ABI_VERSION
        cfg.device, cfg.n_replicates, cfg.replicate_offset = device, n_replicates, replicate_offset
        cfg.n_cap, cfg.t_cap, cfg.compact_every, cfg.seed = n_cap, t_cap, compact_every, seed
        cfg.e_cap[:] = list(e_cap)
        self.cfg = cfg
        self._h = C.c_void_p()
        rc = lib.xgc_create(C.byref(cfg), C.byref(self._h))
        if rc:
            raise XgcError(f"xgc_create failed ({rc}): {lib.xgc_last_error(None).decode()}")
        self.n_replicates, self.n_cap, self.e_cap, self.t_cap = n_replicates, n_cap, list(e_cap), t_cap

    def close(self):
        if getattr(self, "_h", None):
            lib.xgc_destroy(self._h)
            self._h = None

    __del__ = close

    def _ck(self, rc: int):
        if rc:
            raise XgcError(f"xgc error {rc}: {lib.xgc_last_error(self._h).decode()}")

    # configuration -------------------------------------------------------------------------------------------------
    def set_params(self, p: np.ndarray, r: int = -1):
        p = np.ascontiguousarray(p, dtype=np.float64)
        if p.ndim == 2:
            assert p.shape == (self.n_replicates, K.NPARAM)
            self._ck(lib.xgc_set_params_all(self._h, _ptr(p, C.c_double)))
        else:
            assert p.shape == (K.NPARAM,)
            self._ck(lib.xgc_set_params(self._h, r, _ptr(p, C.c_double)))

    def get_params(self, r: int) -> np.ndarray:
        p = np.empty(K.NPARAM)
        self._ck(lib.xgc_get_params(self._h, r, _ptr(p, C.c_double)))
        return p

    def set_mode(self, mode):
        """'strict' (default; bit-exact contract) or 'fast' (replicate-resident kernel, native draws only)."""
        m = {"strict": K.MODE_STRICT, "fast": K.MODE_FAST}.get(mode, mode)
        self._ck(lib.xgc_set_mode(self._h, int(m)))

    def get_mode(self) -> str:
        m = C.c_int32()
        self._ck(lib.xgc_get_mode(self._h, C.byref(m)))
        return "fast" if m.value == K.MODE_FAST else "strict"

    def get_params_all(self) -> np.ndarray:
        p = np.empty((self.n_replicates, K.NPARAM))
        self._ck(lib.xgc_get_params_all(self._h, _ptr(p, C.c_double)))
        return p

    def get_tables(self) -> np.ndarray:
        t = np.empty(K.NTABLE)
        self._ck(lib.xgc_get_tables(self._h, _ptr(t, C.c_double)))
        return t

    def get_control(self, r: int) -> np.ndarray:
        c = np.empty(K.NCONTROL, dtype=np.int32)
        self._ck(lib.xgc_get_control(self._h, r, _ptr(c, C.c_int32)))
        return c

    def set_control(self, c: np.ndarray, r: int = -1):
        c = np.ascontiguousarray(c, dtype=np.int32)
        assert c.shape == (K.NCONTROL,)
        self._ck(lib.xgc_set_control(self._h, r, _ptr(c, C.c_int32)))

    def set_tables(self, t: np.ndarray):
        t = np.ascontiguousarray(t, dtype=np.float64)
        assert t.shape == (K.NTABLE,)
        self._ck(lib.xgc_set_tables(self._h, _ptr(t, C.c_double)))

    # state exchange ------------------------------------------------------------------------------------------------
    def set_population(self, r: int, n: int, at: int = 1):
        self._ck(lib.xgc_set_population(self._h, r, n, at))

    def num_agents(self, r: int) -> int:
        n = C.c_int32()
        self._ck(lib.xgc_num_agents(self._h, r, C.byref(n)))
        return n.value

    def upload_attr(self, r: int, name: str, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        self._ck(lib.xgc_upload_attr(self._h, r, name.encode(), v.ctypes.data_as(C.c_void_p), v.size, K.F64))

    def download_attr(self, r: int, name: str) -> np.ndarray:
        v = np.empty(self.n_cap, dtype=np.float64)
        n = C.c_size_t()
        self._ck(lib.xgc_download_attr(self._h, r, name.encode(), v.ctypes.data_as(C.c_void_p), v.size, K.F64, C.byref(n)))
        return v[: n.value].copy()

    def set_edgelist(self, r: int, layer: int, el: np.ndarray, start: Optional[np.ndarray] = None):
        """el: int array [E, 2] of 1-based R positions (stored column-major for the ABI)."""
        el = np.asarray(el, dtype=np.int32).reshape(-1, 2)
        cm = np.ascontiguousarray(el.T)
        st = None if start is None else np.ascontiguousarray(start, dtype=np.int32)
        self._ck(lib.xgc_set_edgelist(self._h, r, layer, _ptr(cm, C.c_int32), el.shape[0], None if st is None else _ptr(st, C.c_int32)))

    def get_edgelist(self, r: int, layer: int):
        cap = self.e_cap[layer]
        st = np.empty(cap, dtype=np.int32)
        n = C.c_int32()
        flat = np.empty(2 * cap, dtype=np.int32)      # the ABI writes column-major [E,2] with E = n_edges
        rc = lib.xgc_get_edgelist(self._h, r, layer, _ptr(flat, C.c_int32), cap, C.byref(n), _ptr(st, C.c_int32))
        self._ck(rc)
        ne = n.value
        return np.stack([flat[:ne], flat[ne:2 * ne]], axis=1).copy(), st[:ne].copy()

    def get_actcounts(self, r: int, layer: int) -> np.ndarray:
        cap = self.e_cap[layer]
        flat = np.empty(4 * cap, dtype=np.int32)
        n = C.c_int32()
        self._ck(lib.xgc_get_actcounts(self._h, r, layer, _ptr(flat, C.c_int32), cap, C.byref(n)))
        ne = n.value
        return flat[: 4 * ne].reshape(4, ne).T.copy()

    def get_actlist(self, r: int, at: int, cap: int = 1 << 16, inj: Optional[xgc_inject] = None) -> np.ndarray:
        al = np.empty(6 * cap, dtype=np.int32)
        n = C.c_int32()
        self._ck(lib.xgc_get_actlist(self._h, r, at, _ptr(al, C.c_int32), cap, C.byref(n), None if inj is None else C.byref(inj)))
        return al.reshape(6, cap)[:, : n.value].T.copy()

    # stepping ------------------------------------------------------------------------------------------------------
    def step(self, at: int, mask: Optional[int] = None, inj: Optional[xgc_inject] = None):
        self._ck(lib.xgc_step(self._h, at, K.M_ALL if mask is None else mask, None if inj is None else C.byref(inj)))

    def set_replicate_ids(self, ids):
        """Global replicate ids (Philox keys) of this handle's replicates; default replicate_offset + r."""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        assert ids.shape == (self.n_replicates,)
        self._ck(lib.xgc_set_replicate_ids(self._h, _ptr(ids, C.c_int32)))

    def replicate_cost(self, reset: bool = False) -> np.ndarray:
        """SM cycles the resident (fast) kernel spent on each replicate since the last reset."""
        out = np.zeros(self.n_replicates, dtype=np.uint64)
        self._ck(lib.xgc_replicate_cost(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64)), int(reset)))
        return out

    def step_span(self, at: int, mask_tail: int, mask_head: int):
        """The rest of week `at` and the start of week at + 1 in one call (xgc_step_span; one resident launch in fast mode)."""
        self._ck(lib.xgc_step_span(self._h, at, mask_tail, mask_head))

    def week_hostnet(self, at: int, mask_tail: int, mask_head: int, delta_main: int, stride_main: int, delta_casl: int, stride_casl: int,
                     has_start: bool, el_inst: int, ne_inst: int, stride_inst: int, counters_out: int, n_alive_out: int):
        """One week of the host-network loop in one ABI call (xgc_week_hostnet); every buffer is a raw host pointer (0 = NULL)."""
        self._ck(lib.xgc_week_hostnet(self._h, at, mask_tail, mask_head, delta_main or None, stride_main, delta_casl or None, stride_casl,
                                      int(has_start), el_inst or None, ne_inst or None, stride_inst, counters_out or None, n_alive_out or None))

    def run(self, at0: int, at1: int):
        self._ck(lib.xgc_run(self._h, at0, at1))

    def sync(self):
        self._ck(lib.xgc_sync(self._h))

    def compact(self):
        self._ck(lib.xgc_compact(self._h))

    def status(self, r: int) -> int:
        s = C.c_uint32()
        self._ck(lib.xgc_status(self._h, r, C.byref(s)))
        return s.value

    # outputs -------------------------------------------------------------------------------------------------------
    def counters(self, r: int, at0: int, at1: int) -> np.ndarray:
        out = np.empty((at1 - at0 + 1, K.NCOUNTER), dtype=np.uint32)
        self._ck(lib.xgc_get_counters(self._h, r, at0, at1, _ptr(out, C.c_uint32)))
        return out

    def counters_all(self, at: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.n_replicates, K.NCOUNTER), dtype=np.uint32)
        self._ck(lib.xgc_get_counters_all(self._h, at, out.ctypes.data_as(C.c_void_p)))
        return out

    def epi(self, r: int, var: str, at0: int, at1: int) -> np.ndarray:
        out = np.empty(at1 - at0 + 1)
        self._ck(lib.xgc_get_epi(self._h, r, var.encode(), _ptr(out, C.c_double), at0, at1))
        return out

    def epi_table(self, r: int, names, at0: int, at1: int) -> np.ndarray:
        """[len(names), at1 - at0 + 1] series of replicate r with one device->host read (xgc_get_epi_table)."""
        names = list(names)
        arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        out = np.empty((len(names), at1 - at0 + 1))
        self._ck(lib.xgc_get_epi_table(self._h, r, arr, len(names), _ptr(out, C.c_double), at0, at1))
        return out

    def targets(self, at_last: int, tailn: int) -> np.ndarray:
        out = np.empty((self.n_replicates, K.NTARGET))
        self._ck(lib.xgc_targets(self._h, at_last, tailn, _ptr(out, C.c_double)))
        return out

    # checkpoint ----------------------------------------------------------------------------------------------------
    def snapshot(self) -> np.ndarray:
        n = C.c_size_t()
        self._ck(lib.xgc_snapshot_size(self._h, C.byref(n)))
        buf = np.empty(n.value, dtype=np.uint8)
        self._ck(lib.xgc_snapshot(self._h, buf.ctypes.data_as(C.c_void_p), n.value))
        return buf

    def restore(self, buf: np.ndarray):
        self._ck(lib.xgc_restore(self._h, buf.ctypes.data_as(C.c_void_p), buf.size))

    def fork_from(self, src: "Handle", src_r: int, dst_r: int):
        self._ck(lib.xgc_fork(src._h, src_r, self._h, dst_r))

    def set_edgelists_all(self, layer: int, el_ptr: int, ne_ptr: int, start_ptr: int = 0, stride: int = 0):
        """Raw-pointer variant for pinned host buffers: el [R][2][stride] int32 1-based positions, n_edges [R]."""
        self._ck(lib.xgc_set_edgelists_all(self._h, layer, el_ptr, ne_ptr, start_ptr or None, stride or self.e_cap[layer]))

    def update_edgelists_all(self, layer: int, delta_ptr: int, stride: int, has_start: bool = True):
        """Weekly delta of a persistent layer: packed int32 rows [R][stride] = {n_removed, n_added, removed idx (0-based, ascending),
        tails, heads (1-based positions)(, start weeks)}; raw pointer to (ideally pinned) host memory."""
        self._ck(lib.xgc_update_edgelists_all(self._h, layer, delta_ptr, stride, int(has_start)))

    def num_agents_all(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.n_replicates, dtype=np.int32)
        self._ck(lib.xgc_num_agents_all(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def profile(self, enable: bool):
        self._ck(lib.xgc_profile(self._h, int(enable)))

    def profile_read(self) -> dict:
        """{kernel: (total ms, launches)} from CUDA events recorded on the launching stream."""
        cap = 32
        nm, ms, cnt = (C.c_char_p * cap)(), (C.c_double * cap)(), (C.c_uint64 * cap)()
        n = lib.xgc_profile_read(self._h, nm, ms, cnt, cap)
        return {nm[k].decode(): (ms[k], cnt[k]) for k in range(n)}

    def kernel_launches(self) -> int:
        n = C.c_uint64()
        lib.xgc_kernel_launches(self._h, C.byref(n))
        return n.value

    def stream(self) -> int:
        s = C.c_void_p()
        lib.xgc_stream_ptr(self._h, C.byref(s))
        return s.value or 0
