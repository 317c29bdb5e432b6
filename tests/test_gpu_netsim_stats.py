"""GPU tests of the reference-facing host interface (param_msm / control_msm / init_msm / netsim, module(dat, at) -> dat)
and north_star's second correctness test: with native RNG, prevalence / incidence / screening targets across >= 1000
replicates are statistically indistinguishable between the CUDA path and the CPU oracle run with DIFFERENT seeds."""
import numpy as np
import pytest

from common import K, P

pytestmark = pytest.mark.gpu


def test_netsim_mirror_and_module_contract():
    from xgc_msm_b200 import modules as M
    from xgc_msm_b200 import control_msm, init_msm, param_msm
    param = param_msm(u2rgc_tprob=0.2, u2pgc_tprob=0.2, r2ugc_tprob=0.25, p2ugc_tprob=0.1, rgc_ntx_int=12, gc_asympt_visit_prob=0.02,
                      a_rate=0.00049 + 1.285 / 20000)
    init = init_msm(prev_ugc=0.05, prev_rgc=0.05, prev_pgc=0.05)
    x = {"n": 3000, "seed": 5}
    ctl = control_msm(nsims=3, nsteps=30, ncores=1, network_mode="surrogate", stiScreeningProtocol="base", gcUntreatedRecovDist="geom",
                      skip_check=True, tergmLite=True, verbose=False)
    sim = M.netsim(x, param, init, ctl)                                   # fused weeks (xgc_run)
    epi = sim.epi
    assert epi["num"].shape == (30, 3) and np.all(epi["num"][0] == 3000)   # at = 1 row holds the initial prevalence_msm
    assert np.all(epi["prev.gc"][-1] > 0) and not np.array_equal(epi["incid.gc"][:, 0], epi["incid.gc"][:, 1])
    calls = []

    def my_stitx(dat, at):                                                # a user-supplied module forces the module loop
        calls.append(at)
        return M.stitx_msm(dat, at)

    ctl2 = control_msm(nsims=3, nsteps=30, network_mode="surrogate", stitx_FUN=my_stitx, aging_FUN=M.aging_msm)
    sim2 = M.netsim(x, param, init, ctl2)
    assert calls == list(range(2, 31))
    for v in ("num", "incid.gc", "prev.rgc", "i.prev", "prop.ureth.tested", "incid.p2pgc"):
        assert np.array_equal(np.nan_to_num(sim2.epi[v]), np.nan_to_num(epi[v])), v
    # host-side network step (simnet_msm stays on the host): the hook re-wires layer 2 every week through dat.set_el
    seen = []

    def hook(dat, at):
        for s in range(dat.nsims):
            n = dat.h.num_agents(s)
            el = np.stack([np.arange(1, 41), np.arange(n - 39, n + 1)], axis=1)
            dat.set_el(2, el, sim=s)
        seen.append(at)

    ctl3 = control_msm(nsims=2, nsteps=8, network_mode="host")
    sim3 = M.netsim({"n": 1000, "seed": 2}, param, init, ctl3, network_hook=hook)
    assert seen == list(range(2, 9)) and sim3.dat.el(2)[0].shape == (40, 2)
    # restart from the finished object with another screening arm (02.00_epi.r:292-296)
    base = M.netsim({"n": 1500, "seed": 3, "t_cap": 64}, param, init, control_msm(nsims=2, nsteps=20, network_mode="surrogate"), checkpoint=True)
    arm = M.netsim(base, param, init, control_msm(nsims=2, nsteps=40, start=21, network_mode="surrogate", stiScreeningProtocol="universal"))
    assert arm.dat.at == 40 and arm.epi["num.rgc.test"][25:].sum() > base.epi["num.rgc.test"][:20].sum()


def test_statistical_equivalence_1000_replicates():
    """1000 GPU replicates (Philox seed 1971) vs 1000 oracle replicates (Philox seed 424242): two-sample KS at
    alpha = 0.001 per target and |difference of means| <= 5 standard errors.  Same parameters and starting population."""
    from scipy.stats import ks_2samp
    from oracle import cpu_bench
    from xgc_msm_b200 import targets as T, workload
    from xgc_msm_b200.abi import Handle
    R, N, WEEKS, TAIL = 1000, 1200, 60, 26
    kw = dict(lhs=False, n_distinct=1)
    spec = workload.SweepSpec(n_agents=N, seed=1971, **kw)
    h = Handle(R, spec.n_cap, spec.e_cap, WEEKS + 3, seed=1971)
    workload.load_sweep(h, spec)
    h.run(2, WEEKS + 1)
    g = h.targets(WEEKS + 1, TAIL)
    c = cpu_bench.run_targets(N, R, WEEKS, TAIL, 1971, K.M_ALL, kw, philox_seed=424242)
    assert np.isfinite(g).all() and np.isfinite(c).all() and not np.array_equal(g, c)
    names = T.device_target_names()
    report = {}
    for k in ("num", "i.prev", "ir100.pop", "i.prev.dx.inf.W", "cc.vsupp.W", "prop.rect.tested", "prop.ureth.tested", "prop.phar.tested",
              "prob.rGC.tested", "prob.uGC.tested", "prob.pGC.tested", "prev.gc", "prev.rgc", "prev.ugc", "prev.pgc",
              "ir100.gc", "ir100.rgc", "ir100.ugc", "ir100.pgc"):
        j = names.index(k)
        p = ks_2samp(g[:, j], c[:, j]).pvalue
        se = np.sqrt(g[:, j].var(ddof=1) / R + c[:, j].var(ddof=1) / R)
        z = abs(g[:, j].mean() - c[:, j].mean()) / max(se, 1e-300)
        report[k] = (round(float(p), 4), round(float(z), 2))
        assert p > 1e-3 and z < 5.0, (k, report[k], g[:, j].mean(), c[:, j].mean())
    print("KS p-value, |z| per target:", report)
    # and the same seed gives the same numbers exactly (counter-based RNG): one replicate, CPU vs GPU
    c0 = cpu_bench.run_targets(N, 2, WEEKS, TAIL, 1971, K.M_ALL, kw, philox_seed=1971, cores=2)
    assert np.array_equal(c0.view(np.uint64), g[:2].view(np.uint64))


def test_screening_scenario_grid():
    """BASELINE configs[2]: burn-in -> fork into the five screening arms -> NIA / PIA / NTIA (99_analyze.r:146-208)."""
    from xgc_msm_b200 import scenarios, workload
    from xgc_msm_b200.abi import Handle
    from xgc_msm_b200.params import control_msm
    R, N, BURN, CONT = 6, 2500, 26, 30
    spec = workload.SweepSpec(n_agents=N, seed=1971, lhs=False, n_distinct=1)
    h = Handle(R, spec.n_cap, spec.e_cap, BURN + CONT + 4, seed=1971)
    workload.load_sweep(h, spec)
    h.run(2, BURN + 1)
    res = scenarios.run_screening_grid(h, BURN + 1, CONT, control_msm(network_mode="surrogate"))
    assert set(res) == {"base", "symp", "cdc1", "cdc2", "univ"}
    assert np.all(res["base"]["nia_gc"] == 0) and np.all(res["symp"]["asympt_visits"] == 0)
    assert np.all(res["base"]["asympt_visits"] > 0) and np.all(res["cdc1"]["asympt_visits"] > 0)
    assert res["cdc2"]["sum_num.pgc.test"].sum() > 0.9 * res["cdc1"]["sum_num.pgc.test"].sum()   # kissing as an extra pharyngeal exposure
    assert res["univ"]["sum_num.gc.test"].sum() > res["base"]["sum_num.gc.test"].sum()
    assert res["symp"]["sum_num.gc.test"].sum() < res["base"]["sum_num.gc.test"].sum()        # symptomatic-only testing tests least
    for k in ("nia_gc", "pia_rgc", "ntia_pgc", "sum_incid.ugc"):
        assert res["cdc1"][k].shape == (R,)


def test_full_size_sweep_properties():
    """BASELINE configs[1] at full size, ALL 1000 replicates x 10k agents with LHS parameters: size-independent
    properties -- accounting identities of the weekly tallies, run-mode idempotence (CUDA-graph xgc_run == week-by-week
    xgc_step, bit for bit), and shard invariance (two 500-replicate handles with replicate offsets == the 1000-replicate
    handle).  Bit-exact parity against the oracle on sampled replicates of this same sweep (104 weeks, 10k and 20k
    agents, run-graph and host-network paths) is tests/test_gpu_fullsize_parity.py."""
    from xgc_msm_b200 import workload as W
    from xgc_msm_b200.abi import Handle, K
    R, weeks = 1000, 12
    spec = W.SweepSpec(n_agents=10000, seed=1971)

    def make(r0, n):
        h = Handle(n, spec.n_cap, spec.e_cap, weeks + 4, seed=1971, device=0, replicate_offset=r0, compact_every=5)
        W.load_sweep(h, spec, rid0=r0)
        return h

    full = make(0, R)
    full.run(2, 1 + weeks)                                   # graph replay incl. two slot compactions (weeks 5, 10)
    C = np.stack([full.counters_all(t) for t in range(1, 2 + weeks)]).astype(np.int64)       # [week, R, counter]; row 0 = week 1
    G = K.NCELL
    grp = lambda g: C[:, :, g * G:(g + 1) * G].sum(-1)
    num, inum, gc_any = grp(K.K_NUM), grp(K.K_INUM), grp(K.K_INUM_GC)
    site = [grp(K.K_INUM_RGC + s) for s in range(3)]
    incid_site = sum(grp(K.K_INCID_RGC + s) for s in range(3))
    assert full.status(0) == 0 and int(np.bitwise_or.reduce([full.status(r) for r in range(0, R, 37)])) == 0
    assert np.array_equal(num[-1], full.num_agents_all())                                   # tallied population == living slots
    dep_g, dep_a, nnew = C[2:, :, K.K_DEP_GEN], C[2:, :, K.K_DEP_AIDS], C[2:, :, K.K_NNEW]     # week 1 (row 0) holds no tallies
    dnum = num[2:] - num[1:-1]
    assert np.all(dnum <= nnew - np.maximum(dep_g, dep_a)) and np.all(dnum >= nnew - dep_g - dep_a)   # an agent may draw both deaths
    assert np.array_equal(C[1:, :, K.K_PATH:K.K_PATH + K.NPATH].sum(-1), incid_site[1:])    # every new site infection has one pathway
    assert np.all(gc_any <= site[0] + site[1] + site[2]) and np.all(site[0] + site[1] + site[2] <= 3 * gc_any)
    assert np.all(inum[2:] <= inum[1:-1] + grp(K.K_INCID_HIV)[2:]) and np.all(grp(K.K_INUM_DX) <= inum) and np.all(grp(K.K_VSUPP) <= grp(K.K_INUM_DX))
    assert np.all(C[1:, :, K.K_POS:K.K_POS + 3] <= C[1:, :, K.K_TESTED:K.K_TESTED + 3]) and np.all(C[1:, :, K.K_VISITS_SYMPT] <= C[1:, :, K.K_VISITS])
    assert (C[1:, :, K.K_NACTS] > 0).all() and incid_site[1:].sum() > 0 and grp(K.K_INCID_HIV)[1:].sum() > 0
    # the parameters differ per replicate (LHS): so do the epidemics
    assert np.unique(gc_any[-1]).size > 100

    # run-mode idempotence and shard invariance, bit for bit on every tally of every week
    halves = [make(0, R // 2), make(R // 2, R // 2)]
    for at in range(2, 2 + weeks):
        for h in halves:
            h.step(at)
            if at % 5 == 0:
                h.compact()
    for k, h in enumerate(halves):
        for t in range(2, 2 + weeks):
            got, want = h.counters_all(t), C[t - 1, k * (R // 2):(k + 1) * (R // 2)].astype(np.uint32)
            if not np.array_equal(got, want):
                rr, cc = np.nonzero(got != want)
                raise AssertionError(f"shard {k}, week {t}: {np.unique(rr).size} replicates differ (first {np.unique(rr)[:8]}), counters {np.unique(cc)[:16]}, "
                                     f"first: shard {got[rr[0], cc[0]]} vs full handle {want[rr[0], cc[0]]}")
        h.close()
    tg = full.targets(1 + weeks, 5)
    assert tg.shape == (R, K.NTARGET) and np.isfinite(tg[:, :3]).all()
    full.close()
