timeout 600 python -m pytest tests/test_gpu_multilevel.py tests/test_gpu_two_level.py -x -q > gpurun_out/r02_c13_pytest.log 2>&1; tail -3 gpurun_out/r02_c13_pytest.log
bash tools/gpu_call3.sh
