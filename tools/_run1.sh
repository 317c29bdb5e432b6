timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "config5 or large_population or host_network" 2>&1 | tail -2
for R in 1 8; do
python bench.py --replicates $R --agents 1000000 --mode strict --steps 60 --no-cpu-baseline --e2e-steps 3 > gpurun_out/big1m_e$R.json 2> gpurun_out/big1m_e$R.err
tail -2 gpurun_out/big1m_e$R.err
python -c "
import json; d=json.loads(open('gpurun_out/big1m_e$R.json').read()); print(d['value'], d['ms_per_step'], d['roofline']['kernels_ms_per_step']['k_rank'], d['roofline']['step'])"
done
python tools/run_burnin.py > gpurun_out/burnin_3120_r02.json 2> gpurun_out/burnin_r02.err; tail -3 gpurun_out/burnin_r02.err; cat gpurun_out/burnin_3120_r02.json | head -c 1500
