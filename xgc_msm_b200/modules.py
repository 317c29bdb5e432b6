"""Host-side mirror of the reference's module interface: `FUN(dat, at) -> dat`, `initialize_msm`, `reinit_msm`, `netsim`.

Reference call sites: /root/reference/burnin/cal/sim1_02_lhs.r:195-232 (control_msm(<x>.FUN = ...), netsim(est, param,
init, control)) and inst/analysis02_screening/02.00_epi.r:289-329 (restart with reinit_msm, start = 3121).
Each module function launches the CUDA kernels of that module through the C ABI (xgc_step with one mask bit); the
state (`dat$attr`, `dat$el`, `dat$epi`) stays on the device behind `dat.h` and is materialised only when asked for.
R is not installed in this image, so this Python mirror stands where R/modules.R + src/xgc_r.c would.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional

import numpy as np

from . import params as P
from . import synth
from .abi import Handle, K, names


@dataclass
class Dat:
    """The `dat` object of EpiModel: device state handle + the R-visible pieces."""
    h: Handle
    param: np.ndarray            # [nsims, NPARAM]
    control: P.Control
    nsims: int
    at: int = 1
    network_hook: Optional[Callable[["Dat", int], None]] = None     # host-side simnet_msm (network_mode "host")

    # dat$attr / dat$el / dat$epi views ---------------------------------------------------------------------------
    def attr(self, name: str, sim: int = 0) -> np.ndarray:
        return self.h.download_attr(sim, name)

    def set_attr(self, name: str, v, sim: int = 0):
        self.h.upload_attr(sim, name, v)

    def el(self, layer: int, sim: int = 0):
        return self.h.get_edgelist(sim, layer)

    def set_el(self, layer: int, el, start=None, sim: int = 0):
        self.h.set_edgelist(sim, layer, el, start)

    def epi(self, var: str, sim: int = 0, at0: int = 1, at1: Optional[int] = None) -> np.ndarray:
        return self.h.epi(sim, var, at0, self.at if at1 is None else at1)


def _module(name: str):
    bit = P.MODULE_BIT[name]

    def fun(dat: Dat, at: int) -> Dat:
        dat.h.step(at, bit)
        dat.at = at
        return dat

    fun.__name__ = f"{name}_msm"
    fun.__doc__ = f"{name}_msm(dat, at) -> dat: runs the '{name}' module of week `at` on the device (mask bit {bit:#x})."
    fun.xgc_bit = bit
    return fun


aging_msm = _module("aging")            # sim1_02_lhs.r:202
departure_msm = _module("departure")    # :203
arrival_msm = _module("arrival")        # :204
hivtest_msm = _module("hivtest")        # :205
hivtx_msm = _module("hivtx")            # :206
hivprogress_msm = _module("hivprogress")  # :207
hivvl_msm = _module("hivvl")            # :208
acts_msm = _module("acts")              # :210
condoms_msm = _module("condoms")        # :211
position_msm = _module("position")      # :212
prep_msm = _module("prep")              # :213
hivtrans_msm = _module("hivtrans")      # :214
stirecov_msm = _module("stirecov")      # :215
stitx_msm = _module("stitx")            # :216
stitrans_msm_rand = _module("stitrans")  # :217
prevalence_msm = _module("prev")        # :218
_resim_surrogate = _module("resim_nets")


def simnet_msm(dat: Dat, at: int) -> Dat:
    """resim_nets.FUN (sim1_02_lhs.r:209).  network_mode 'host': the STERGM resimulation stays on the host — the hook
    receives `dat`, reads what it needs (dat.attr('race') ..., dat.h.num_agents) and writes dat.set_el(layer, el, start).
    network_mode 'surrogate': on-device dissolution/formation stand-in (benchmarks, statistical tests)."""
    if dat.control.flags[P.CONTROL_INDEX["network_mode"]] == 1:
        return _resim_surrogate(dat, at)
    if dat.network_hook is not None:
        dat.network_hook(dat, at)
    dat.at = at
    return dat


simnet_msm.xgc_bit = P.MODULE_BIT["resim_nets"]

DEFAULT_MODULES: Dict[str, Callable] = {
    "aging": aging_msm, "departure": departure_msm, "arrival": arrival_msm, "hivtest": hivtest_msm, "hivtx": hivtx_msm,
    "hivprogress": hivprogress_msm, "hivvl": hivvl_msm, "resim_nets": simnet_msm, "acts": acts_msm, "condoms": condoms_msm,
    "position": position_msm, "prep": prep_msm, "hivtrans": hivtrans_msm, "stirecov": stirecov_msm, "stitx": stitx_msm,
    "stitrans": stitrans_msm_rand, "prev": prevalence_msm}


def initialize_msm(x: dict, param, init: dict, control: P.Control, s: int = 0, device: int = 0, seed: int = 1971,
                   replicate_offset: int = 0) -> Dat:
    """initialize.FUN(x, param, init, control, s) (sim1_02_lhs.r:201).  `x` stands for the fitted netest object:
    {"n": agents, "seed": int} for a synthetic population, optionally "attrs"/"els"/"starts"/"tables" to supply them."""
    nsims = control.nsims
    param = np.atleast_2d(np.asarray(param, dtype=np.float64))
    if param.shape[0] == 1 and nsims > 1:
        param = np.repeat(param, nsims, axis=0)
    n = int(x["n"])
    n_cap = x.get("n_cap") or synth.capacity(n)
    e_cap = x.get("e_cap") or synth.edge_capacity(n)
    h = Handle(nsims, n_cap, e_cap, int(x.get("t_cap", control.nsteps + 2)), seed=seed, device=device, replicate_offset=replicate_offset)
    h.set_tables(x.get("tables", synth.make_tables()))
    p0 = param[0].copy()
    attrs = x.get("attrs") or synth.init_population(n, int(x.get("seed", 0)), init, p0)
    if "els" in x:
        els, starts = x["els"], x.get("starts", [None] * 3)
    else:
        els, starts = synth.init_network(attrs, int(x.get("seed", 0)))
    h.set_population(0, n, 1)
    for k, v in attrs.items():
        h.upload_attr(0, k, v)
    for l in range(3):
        h.set_edgelist(0, l, els[l], starts[l])
    for r in range(1, nsims):
        h.fork_from(h, 0, r)
    param[:, P.PARAM_INDEX["n_init"][0]] = n
    h.set_params(param if nsims > 1 else param[0])
    h.set_control(control.flags)
    dat = Dat(h=h, param=param, control=control, nsims=nsims, at=1)
    prev = control.modules.get("prev", prevalence_msm)
    return prev(dat, 1)                                          # prevalence_msm(dat, at = 1) [BK]


def reinit_msm(x: "Sim", param, init: dict, control: P.Control, s: int = 0) -> Dat:
    """Restart from a finished simulation object (02.00_epi.r:40-76, 292-296): the saved state becomes the start state,
    parameters/control flags may change (screening arms), `control.start` is the first week to simulate."""
    h = x.dat.h
    if control.nsteps + 2 > h.t_cap:
        raise ValueError(f"reinit_msm: nsteps {control.nsteps} exceeds the handle's t_cap {h.t_cap}; create the burn-in with room for the continuation")
    if x.checkpoint is not None:
        h.restore(x.checkpoint)
    param = np.atleast_2d(np.asarray(param, dtype=np.float64))
    if param.shape[0] == 1 and x.dat.nsims > 1:
        param = np.repeat(param, x.dat.nsims, axis=0)
    h.set_params(param if x.dat.nsims > 1 else param[0])
    h.set_control(control.flags)
    return Dat(h=h, param=param, control=control, nsims=x.dat.nsims, at=control.start - 1, network_hook=x.dat.network_hook)


@dataclass
class Sim:
    """What netsim returns (saveout.net): epi time series, final attributes / edge lists on demand, a checkpoint."""
    dat: Dat
    control: P.Control
    checkpoint: Optional[np.ndarray] = None

    @property
    def epi(self) -> Dict[str, np.ndarray]:
        T = self.dat.at
        return {v: np.stack([self.dat.h.epi(s, v, 1, T) for s in range(self.dat.nsims)], axis=1) for v in names("epi")}

    def targets(self, tailn: int = 52) -> np.ndarray:
        return self.dat.h.targets(self.dat.at, tailn)


def netsim(x, param, init: dict, control: P.Control, device: int = 0, seed: int = 1971, checkpoint: bool = False,
           network_hook=None, mode: Optional[str] = None) -> Sim:
    """EpiModel::netsim: dat <- initialize.FUN(...); for at in max(2, start):nsteps, for each module: dat <- FUN(dat, at)
    (sim1_02_lhs.r:232; SURVEY Appendix E "Loop/order").  When every module is the library's own, consecutive modules
    are issued as ONE fused xgc_step per week (identical results: test_module_by_module_contract)."""
    mods = [control.modules.get(m, DEFAULT_MODULES[m]) for m in P.MODULE_ORDER]
    init_fun = control.modules.get("initialize")
    if isinstance(x, Sim):
        dat = (init_fun or reinit_msm)(x, param, init, control, 0)
    else:
        dat = (init_fun or (lambda *a: initialize_msm(*a, device=device, seed=seed)))(x, param, init, control, 0)
    if network_hook is not None:
        dat.network_hook = network_hook
    if mode is not None:                 # "fast": replicate-resident native-draw kernels (statistical contract); "strict": bit-exact
        dat.h.set_mode(mode)
    first = max(2, control.start)
    all_native = all(getattr(f, "xgc_bit", None) is not None for f in mods)
    host_net = control.flags[P.CONTROL_INDEX["network_mode"]] == 0 and dat.network_hook is not None
    if all_native and not host_net:
        if first <= control.nsteps:
            dat.h.run(first, control.nsteps)
            dat.at = control.nsteps
    else:
        i = P.MODULE_ORDER.index("resim_nets")
        pre_mask, post_mask = sum(P.MODULE_BIT[m] for m in P.MODULE_ORDER[:i]), sum(P.MODULE_BIT[m] for m in P.MODULE_ORDER[i + 1:])
        pre_done = -1
        for at in range(first, control.nsteps + 1):
            if all_native:          # fuse the native modules either side of the host network step
                if pre_done != at:
                    dat.h.step(at, pre_mask)
                dat = simnet_msm(dat, at)
                if at < control.nsteps:     # nothing on the host between prevalence_msm(at) and aging_msm(at + 1): one call (xgc_step_span)
                    dat.h.step_span(at, post_mask, pre_mask); pre_done = at + 1
                else:
                    dat.h.step(at, post_mask)
                dat.at = at
            else:
                for f in mods:
                    dat = f(dat, at)
    dat.h.sync()
    return Sim(dat=dat, control=control, checkpoint=dat.h.snapshot() if checkpoint else None)
