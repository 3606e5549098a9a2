timeout 900 python tests/stokes_bench.py 224 352 8 0 5 500 0 poisson 1e-8 2>&1 | tail -1 | grep -o '"gmres.*'
timeout 900 python tests/stokes_bench.py 224 352 8 0 5 500 0 poisson 1e-10 2>&1 | tail -1 | grep -o '"gmres.*'
