timeout 300 python tools/e2e_zero_copy.py > gpurun_out/r02_c59_zero_copy.log 2>&1; echo "rc=$?"; cat gpurun_out/r02_c59_zero_copy.log
