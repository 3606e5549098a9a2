timeout 1200 python -m pytest tests/test_gpu_multilevel.py tests/test_gpu_two_level.py tests/test_gpu_condensed.py -x -q > gpurun_out/r02_c38_pytest.log 2>&1; tail -4 gpurun_out/r02_c38_pytest.log
timeout 300 python tests/ml_profile.py 1024 8 4096 2>&1 | tail -2
SEMK_NO_GRAPH=1 timeout 300 python tests/ml_profile.py 1024 8 4096 2>&1 | tail -2
timeout 900 python tests/stokes_bench.py 224 352 8 0 5 500 0 poisson 1e-8 2>&1 | tail -1 | grep -o '"gmres.*'
