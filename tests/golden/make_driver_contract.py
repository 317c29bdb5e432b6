"""Writes tests/golden/driver_contract.json by PARSING the reference's own driver scripts (run in a container where
/root/reference exists; the fixture travels, the reference does not):

  burnin/cal/sim1_02_lhs.r:58-178    argument names of the param_msm(...) call and the env var each one reads
  burnin/cal/sim1_02_lhs.r:180-184   init_msm(...) initial site prevalences
  burnin/cal/sim1_02_lhs.r:195-230   control_msm(...): module slots in the order written, option flags
  inst/analysis02_screening/01_setup.r:116-159   screening arms (scenario, kiss_exposure)
  inst/cal/sim3_01_analyze.r:169-176  the variable patterns summed into `mase`; sim1_01_lhs_params.r:118-120 LHS seed / size

Only names, literals and ordering are recorded (no code).  tests/test_oracle_cpu.py checks the host-side mirror
(params.def, MODULE_ORDER, control_msm defaults, scenarios.ARMS, PRIORS) against it.
"""
import json, os, re

REF = "/root/reference"
src = open(os.path.join(REF, "burnin/cal/sim1_02_lhs.r")).read().split("\n")


def block(start_pat):
    i = next(k for k, l in enumerate(src) if re.search(start_pat, l))
    depth, out = 0, []
    for k in range(i, len(src)):
        line = src[k].split("#")[0]
        depth += line.count("(") - line.count(")")
        out.append((k + 1, src[k]))
        if depth == 0 and k > i:
            break
    return out


params = []
for ln, l in block(r"^param <- param_msm\("):
    m = re.match(r"^\s{2}([A-Za-z0-9_.]+)\s*=\s*(.*)$", l)
    if m:
        env = re.findall(r'fmt_getenv\("([A-Z0-9_]+)"\)', m.group(2))
        params.append({"name": m.group(1), "line": ln, "env": env})
# continuation lines of multi-line c(...) arguments carry further env vars: attach them to the argument above
cur = None
for ln, l in block(r"^param <- param_msm\("):
    m = re.match(r"^\s{2}([A-Za-z0-9_.]+)\s*=", l)
    if m:
        cur = next(p for p in params if p["line"] == ln)
    elif cur is not None:
        for e in re.findall(r'fmt_getenv\("([A-Z0-9_]+)"\)', l):
            if e not in cur["env"]:
                cur["env"].append(e)

init = {m.group(1): float(m.group(2)) for ln, l in block(r"^init <- init_msm\(") for m in [re.match(r"^\s+([a-z.]+)\s*=\s*([0-9.]+)", l)] if m}

modules, flags = [], {}
for ln, l in block(r"^control <- control_msm\("):
    m = re.match(r"^\s+([A-Za-z_.]+)\.FUN\s*=\s*([A-Za-z_.]+)", l)
    if m:
        modules.append({"slot": m.group(1), "fun": m.group(2), "line": ln})
        continue
    m = re.match(r'^\s+([A-Za-z_.]+)\s*=\s*("?)([A-Za-z]+)\2\s*,?', l)
    if m and m.group(1) not in ("nsims", "nsteps", "ncores"):
        flags[m.group(1)] = {"TRUE": True, "FALSE": False}.get(m.group(3), m.group(3))

arms_src = open(os.path.join(REF, "inst/analysis02_screening/01_setup.r")).read()
arms = [{"name": a, "runtype": rt, "scenario": sc, "kiss_exposure": ke == "TRUE"}
        for a, rt, sc, ke in re.findall(r'(sti_[a-z0-9]+) = data\.table\(\s*runtype\s*=\s*"([a-z]+)",\s*scenario\s*=\s*"([a-z]+)",\s*kiss_exposure\s*=\s*"([A-Z]+)"', arms_src)]

an = open(os.path.join(REF, "inst/cal/sim3_01_analyze.r")).read()
vp = re.search(r"varpats <- paste0\(c\((.*?)\), collapse", an, re.S).group(1)
varpats = re.findall(r'"([^"]+)"', vp)
lhs = open(os.path.join(REF, "burnin/cal/sim1_01_lhs_params.r")).read()
lhs_seed = int(re.search(r"set\.seed\((\d+)\)", lhs).group(1))
lhs_n = int(re.search(r"randomLHS\((\d+),", lhs).group(1))
setup = open(os.path.join(REF, "inst/cal/sim0.0_setup.r")).read()
pop = re.search(r"pop_N\s*<-\s*(\d+)", setup)

out = {"mase_varpats": varpats, "lhs_seed": lhs_seed, "lhs_n": lhs_n, "pop_N": int(pop.group(1)) if pop else None,
       "source": "parsed from burnin/cal/sim1_02_lhs.r and inst/analysis02_screening/01_setup.r by tests/golden/make_driver_contract.py",
       "param_msm": params, "init_msm": init, "modules": modules, "control_flags": flags, "screening_arms": arms}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "driver_contract.json"), "w"), indent=1)
print(len(params), "param_msm arguments;", len(modules), "module slots;", flags, ";", [a["name"] for a in arms], init)
