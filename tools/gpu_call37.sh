timeout 1500 python tests/stokes_bench.py 224 352 8 1.0 5 0 0 poisson 1e-8 L newton > gpurun_out/r02_c37_newton.txt 2>&1; tail -8 gpurun_out/r02_c37_newton.txt | cut -c1-1200
