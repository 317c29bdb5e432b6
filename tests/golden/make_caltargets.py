"""Writes tests/golden/caltargets.json: the calibration targets of est/caltargets.Rds.

est/caltargets.Rds is a git-LFS pointer and R is not installed, so the values cannot be produced by running the
reference.  They are the numbers SURVEY.md Appendix D records (recomputed there from the literals in
/root/reference/inst/cal/calval_targets.r:23-322) plus the confidence limits that file states; this script copies them
into a fixture, independent of xgc_msm_b200/targets.py which re-derives them from the raw counts.
"""
import json, os
rows = []
def add(t, vals, subs, ll=None, ul=None):
    for i, (v, s) in enumerate(zip(vals, subs)):
        rows.append({"target": t, "subgroup": s, "value": v, "ll95": None if ll is None else ll[i], "ul95": None if ul is None else ul[i]})
R, A = ["B", "H", "O", "W"], [f"age{k}" for k in range(1, 6)]
add("ct_hiv_prev", [0.124], [None], [0.109], [0.138])                                   # calval_targets.r:212-218
add("ct_hiv_incid_per100pop", [0.514], [None], [0.444], [0.584])                       # :203-209
add("ct_hivdx_pr_byrace", [0.8174, 0.7974, 0.8463, 0.8931], R)                          # :23-75
add("ct_hivdx_pr_byage", [0.5422, 0.7001, 0.8348, 0.9241, 0.9584], A)                   # :77-90
add("ct_vls_pr_byrace", [0.5224, 0.6068, 0.6655, 0.6728], R)                            # :101-137 (Wald limits checked loosely)
add("ct_vls_pr_byage", [0.5088, 0.5515, 0.5978, 0.6385, 0.6545], A)                     # :139-156
add("ct_prep", [0.2624, 0.2997, 0.3983, 0.4237], R, [0.233, 0.274, 0.346, 0.400], [0.293, 0.327, 0.452, 0.448])   # :169-186
add("ct_prop_anatsite_tested", [0.5316, 0.6652, 0.6461], ["rect", "ureth", "phar"], [0.4706, 0.5889, 0.5720], [0.5933, 0.7425, 0.7211])  # :292-308
add("ct_prop_anatsite_pos", [0.118, 0.075, 0.091], ["rGC", "uGC", "pGC"], [0.104, 0.057, 0.079], [0.132, 0.093, 0.103])                  # :313-322
json.dump({"source": "SURVEY.md Appendix D / inst/cal/calval_targets.r literals", "rows": rows},
          open(os.path.join(os.path.dirname(__file__), "caltargets.json"), "w"), indent=1)
