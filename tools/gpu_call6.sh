timeout 900 python -m pytest tests/test_gpu_multilevel.py tests/test_gpu_two_level.py -x -q > gpurun_out/r02_c6_pytest.log 2>&1; tail -8 gpurun_out/r02_c6_pytest.log
timeout 300 python tests/ml_profile.py 1024 8 4096 > gpurun_out/r02_c6_ml4096.txt 2>&1; cat gpurun_out/r02_c6_ml4096.txt
timeout 300 python tests/ml_profile.py 1024 8 2048 > gpurun_out/r02_c6_ml2048.txt 2>&1; cat gpurun_out/r02_c6_ml2048.txt
timeout 300 python tests/ml_profile.py 884 8 4096 > gpurun_out/r02_c6_ml884.txt 2>&1; cat gpurun_out/r02_c6_ml884.txt
