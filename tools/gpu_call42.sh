timeout 1700 python -m pytest tests -x -q -m gpu > gpurun_out/r02_c42_pytest.log 2>&1; tail -5 gpurun_out/r02_c42_pytest.log
