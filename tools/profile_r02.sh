#!/bin/bash
# Round-2 evidence set on ONE B200 (run under gpurun): bench lines, launch list, ncu --set full of the dominant kernel.
O=gpurun_out
python bench.py > $O/bench_r02_default.json 2> $O/bench_r02_default.err || exit 1
python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_r02_reference.json 2> $O/bench_r02_reference.err
python bench.py --mode strict --no-cpu-baseline > $O/bench_r02_strict.json 2> $O/bench_r02_strict.err
python bench.py --agents 20000 --mode strict --no-cpu-baseline --steps 130 > $O/bench_r02_20k_strict.json 2> $O/bench_r02_20k_strict.err
python bench.py --agents 20000 --impl reference --steps 10 --warmup 3 > $O/bench_r02_20k_reference.json 2> $O/bench_r02_20k_reference.err
# launch list of the default command (short run; per-launch times are cold-cache and serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r02.csv \
    python bench.py --steps 64 --warmup 3 --no-cpu-baseline --e2e-steps 2 > $O/ncu_launches_r02.log 2>&1
# --set full of k_week_fast: the warm-up launch (3 weeks) and the timed launch (6 weeks) of a short run
ncu --set full --clock-control none --import-source on -k regex:k_week_fast -c 2 -f -o $O/prof_r02_fast \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 --profile-steps 2 > $O/ncu_full_r02.log 2>&1
tail -2 $O/ncu_full_r02.log
for f in default strict 20k_strict reference 20k_reference; do python - $O/bench_r02_$f.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.3e" % d["value"], "ms/step %.4f" % d["ms_per_step"], "e2e %.3e" % d["e2e"]["value"], (d.get("roofline") or {}).get("frac"))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
done
