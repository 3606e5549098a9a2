timeout 900 python -m pytest tests/test_gpu_multilevel.py tests/test_gpu_two_level.py -x -q > gpurun_out/r02_c2_pytest.log 2>&1; tail -15 gpurun_out/r02_c2_pytest.log
timeout 400 python tests/two_level_bench.py 1024 8 1e-2 three-level > gpurun_out/r02_c2_tl.txt 2>&1; tail -6 gpurun_out/r02_c2_tl.txt
timeout 300 python tests/two_level_bench.py 1024 8 1e-2 three-level 1024 skip-two-level > gpurun_out/r02_c2_tl1024.txt 2>&1; tail -3 gpurun_out/r02_c2_tl1024.txt
timeout 600 python bench.py > gpurun_out/r02_c2_bench.json 2> gpurun_out/r02_c2_bench.err; tail -c 3000 gpurun_out/r02_c2_bench.json; tail -5 gpurun_out/r02_c2_bench.err
