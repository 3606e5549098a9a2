timeout 900 python -m pytest tests/test_gpu_stokes.py -x -q > gpurun_out/r02_c41_pytest.log 2>&1; tail -4 gpurun_out/r02_c41_pytest.log
timeout 1500 python examples/squirmer_axisymmetric.py --swim --nr 15 --nt 9 --order 8 --r-out 100 --re 1 --beta 1 > gpurun_out/r02_c41_swim.txt 2>&1; grep -v "Iteration" gpurun_out/r02_c41_swim.txt | tail -12
timeout 900 python examples/squirmer_axisymmetric.py --swim --nr 15 --nt 9 --order 8 --r-out 100 --re 0 --beta 1 2>&1 | grep -v Iteration | tail -4
