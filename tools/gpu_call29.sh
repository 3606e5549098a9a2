timeout 900 python -m pytest tests/test_gpu_stokes.py tests/test_gpu_condensed.py -x -q > gpurun_out/r02_c29_pytest.log 2>&1; tail -4 gpurun_out/r02_c29_pytest.log
timeout 600 python tests/stokes_bench.py 48 64 8 0 5 400 0 poisson 1e-6 2>&1 | tail -1 | grep -o '"gmres.*'
timeout 600 python tests/stokes_bench.py 48 64 8 0 5 400 0 poisson 1e-4 2>&1 | tail -1 | grep -o '"gmres.*'
timeout 900 python tests/stokes_bench.py 224 352 8 0 5 400 0 poisson 1e-6 2>&1 | tail -1 | grep -o '"gmres.*'
