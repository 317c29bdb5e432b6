python -m pytest tests/test_gpu_fast_mode.py -x -q -k "structure or step_groups" 2>&1 | tail -5
python tools/time_stepgroups.py --replicates 1000
python tools/time_modes.py --modes fast --weeks 128
python -m pytest tests/test_gpu_parity.py -x -q -k config5 2>&1 | tail -5
python bench.py --replicates 1 --agents 1000000 --mode strict --steps 60 --no-cpu-baseline --e2e-steps 3 > gpurun_out/big1m_a.json 2> gpurun_out/big1m_a.err
tail -2 gpurun_out/big1m_a.err
python -c "
import json; d=json.loads(open('gpurun_out/big1m_a.json').read()); print(d['value'], d['ms_per_step'], d['roofline']['kernels_ms_per_step'], d['roofline']['step'])"
