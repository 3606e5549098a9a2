python tools/e2e_timeline.py 16 > gpurun_out/r02_c51_timeline.log 2>&1; cat gpurun_out/r02_c51_timeline.log
