timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_c8_pytest.log 2>&1; tail -8 gpurun_out/r02_c8_pytest.log
