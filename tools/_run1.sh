python tools/run_burnin.py 1000 3120 fast r02 2> gpurun_out/burnin_r02.err | head -c 1600; tail -2 gpurun_out/burnin_r02.err
