"""Calibration round and epi export on the device: LHS -> sweep -> targets -> selection -> narrowed
priors, and the sim$epi-style long table."""
import os

import numpy as np
import pytest

from xgc_msm_b200 import calibration as CAL
from xgc_msm_b200 import params as P
from xgc_msm_b200 import targets as T

pytestmark = pytest.mark.gpu


def test_calibration_round_and_narrowing():
    n, weeks, tailn = 96, 40, 20
    res = CAL.run_round(P.PRIORS, n, weeks, tailn=tailn, n_agents=2000, seed=11, pop_tol=0.10)
    assert res.status == 0
    assert res.inputs.shape == (n, 62) and res.targets.shape == (n, len(T.device_target_names()))
    num = res.targets[:, res.names.index("num")]
    assert np.all(num > 1500) and np.all(num < 2500)
    assert res.keep.sum() >= n // 2 and res.mase is not None and res.mase.shape == (res.keep.sum(),)
    assert np.all(np.isfinite(res.mase)) and np.all(res.mase >= 0)
    # the parameters reached the device: prevalence responds to the LHS rows (not one constant column)
    assert np.std(res.targets[:, res.names.index("prev.gc")]) > 0
    new = CAL.narrow(res, q=0.25)
    changed = 0
    for (n0, lo0, hi0), (n1, lo1, hi1) in zip(P.PRIORS, new):
        assert n0 == n1 and lo0 <= lo1 <= hi1 <= hi0
        changed += (lo1, hi1) != (lo0, hi0)
    assert changed >= 50
    # same seed -> same round, bit for bit (Philox keyed by (seed, replicate id))
    res2 = CAL.run_round(P.PRIORS, n, weeks, tailn=tailn, n_agents=2000, seed=11, pop_tol=0.10)
    assert np.array_equal(res.targets, res2.targets, equal_nan=True)


def test_epi_frame_matches_series_and_round_trips(tmp_path):
    from xgc_msm_b200 import workload as W
    from xgc_msm_b200.abi import Handle, K
    from xgc_msm_b200.export import KEEPVARS, epi_frame, read_epi, write_epi
    spec = W.SweepSpec(n_agents=1500, seed=5)
    h = Handle(6, spec.n_cap, spec.e_cap, 40, seed=5, device=0, replicate_offset=100, compact_every=spec.weeks_between_compactions)
    W.load_sweep(h, spec, rid0=100)
    h.run(2, 30)
    fr = epi_frame(h, 2, 30, names=KEEPVARS, first_time=10)
    T_ = 21
    assert fr["simid"].shape == (6 * T_,) and fr["simid"][0] == 101 and fr["simid"][-1] == 106
    assert np.array_equal(fr["at"][:T_], np.arange(10, 31))
    for r, name in ((0, "num"), (3, "i.prev.B"), (5, "prob.uGC.tested"), (2, "cc.vsupp.age3")):
        assert np.array_equal(fr[name][r * T_:(r + 1) * T_], h.epi(r, name, 10, 30), equal_nan=True)
    G = K.NCELL
    c = h.counters_all(17)
    assert fr["num"][2 * T_ + 7] == float(c[2, K.K_NUM * G:(K.K_NUM + 1) * G].sum())
    full = epi_frame(h, 29, 30, replicates=[1])          # every series the library serves
    assert len(full) == 2 + 193 and full["at"].tolist() == [29, 30]
    with pytest.raises(Exception):
        h.epi_table(0, ["no.such.series"], 2, 3)
    for ext in ("parquet", "feather"):
        p = write_epi(os.path.join(tmp_path, f"epi.{ext}"), fr)
        back = read_epi(p)
        assert list(back) == list(fr)
        for k in fr:
            assert np.array_equal(back[k], fr[k], equal_nan=True)
    h.close()
