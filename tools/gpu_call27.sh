timeout 900 python -m pytest tests/test_gpu_stokes.py -x -q > gpurun_out/r02_c27_pytest.log 2>&1; tail -5 gpurun_out/r02_c27_pytest.log
timeout 600 python tests/stokes_bench.py 48 64 8 0 5 400 0 poisson 2>&1 | tail -1 | grep -o '"gmres.*'
timeout 900 python tests/stokes_bench.py 224 352 8 0 5 400 0 poisson > gpurun_out/r02_c27_stokes_solve.txt 2>&1; tail -1 gpurun_out/r02_c27_stokes_solve.txt | grep -o '"setup_seconds.*' | cut -c1-900
