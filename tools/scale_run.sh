#!/bin/bash
# One multi-GPU measurement set (run under `gpurun --gpus N`): the weak-scaling line the driver runs, the strong-scaling line of
# the north_star configuration (the 1000-replicate sweep split over the GPUs) in both modes, and the reference arm.
#   tools/scale_run.sh N [tag]
N=${1:-8}; TAG=${2:-r02}; OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
[ "$N" = 1 ] && TR="python"
$TR bench.py --gpus $N --no-cpu-baseline                                   > $OUT/scale_${TAG}_weak_fast_$N.json   2> $OUT/scale_${TAG}_$N.err
$TR bench.py --gpus $N --no-cpu-baseline --total-replicates 1000           > $OUT/scale_${TAG}_strong_fast_$N.json 2>> $OUT/scale_${TAG}_$N.err
$TR bench.py --gpus $N --no-cpu-baseline --total-replicates 1000 --mode strict > $OUT/scale_${TAG}_strong_strict_$N.json 2>> $OUT/scale_${TAG}_$N.err
$TR bench.py --gpus $N --impl reference --steps 20 --warmup 3              > $OUT/scale_${TAG}_reference_$N.json   2>> $OUT/scale_${TAG}_$N.err
tail -c 400 $OUT/scale_${TAG}_$N.err
for f in weak_fast strong_fast strong_strict reference; do python - "$OUT/scale_${TAG}_${f}_$N.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.3e" % d["value"], "ms/step %.4f" % d["ms_per_step"], "e2e %.3e" % d["e2e"]["value"], d["config"].get("mode"), d.get("scaling"))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
done
