for tol in 1e-8 1e-6; do
timeout 600 python tests/stokes_bench.py 48 64 8 0 5 400 0 poisson $tol 2>&1 | tail -1 | grep -o '"gmres.*'
done
timeout 900 python tests/stokes_bench.py 224 352 8 0 5 400 0 poisson 1e-8 2>&1 | tail -1 | grep -o '"gmres.*'
