timeout 900 python -m pytest tests/test_gpu_multilevel.py -x -q -k "overlapped or halo_exchange" > gpurun_out/r02_c23_pytest.log 2>&1; tail -15 gpurun_out/r02_c23_pytest.log
timeout 1700 python -m pytest tests -x -q -m gpu > gpurun_out/r02_c23_pytest_full.log 2>&1; tail -8 gpurun_out/r02_c23_pytest_full.log
