timeout 900 python -m pytest tests/test_gpu_stokes.py -x -q > gpurun_out/r02_c20_pytest.log 2>&1; tail -25 gpurun_out/r02_c20_pytest.log
timeout 300 python tests/stokes_bench.py 224 352 8 0 50 > gpurun_out/r02_c20_stokes_bench.txt 2>&1; tail -3 gpurun_out/r02_c20_stokes_bench.txt
timeout 300 python tests/stokes_bench.py 224 352 8 1.0 50 > gpurun_out/r02_c20_stokes_bench_adv.txt 2>&1; tail -3 gpurun_out/r02_c20_stokes_bench_adv.txt
