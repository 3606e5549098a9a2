python tests/ml_profile.py > gpurun_out/r02_c5_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 260 --csv --log-file gpurun_out/r02_c5_launches.csv python tests/ml_profile.py > gpurun_out/r02_c5_ncu.log 2>&1
cat gpurun_out/r02_c5_plain.log; tail -3 gpurun_out/r02_c5_ncu.log
