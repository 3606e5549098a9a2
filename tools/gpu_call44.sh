timeout 1200 python -m pytest tests/test_gpu_condensed.py tests/test_gpu_stokes.py tests/test_gpu_multilevel.py -x -q > gpurun_out/r02_c44_pytest.log 2>&1; tail -6 gpurun_out/r02_c44_pytest.log
timeout 600 python tests/stokes_prec_profile.py 2>&1 | tail -2
timeout 900 python tests/stokes_bench.py 224 352 8 0 5 500 0 poisson 1e-8 2>&1 | tail -1 | grep -o '"gmres.*'
