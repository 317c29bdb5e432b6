for G in 7 10 14; do
python bench.py --no-cpu-baseline --steps 64 --e2e-groups $G --e2e-steps 24 > gpurun_out/e2e_g$G.json 2> gpurun_out/e2e_g$G.err
python -c "
import json; d=json.loads(open('gpurun_out/e2e_g$G.json').read()); print($G, d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])"
done
