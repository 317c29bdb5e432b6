python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --no-cpu-baseline --steps 64 --e2e-steps 24 > gpurun_out/e2e_i.json 2> gpurun_out/e2e_i.err
python -c "
import json; d=json.loads(open('gpurun_out/e2e_i.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])"
