timeout 400 python tests/two_level_bench.py 1024 8 1e-2 three-level > gpurun_out/r02_three_level_first.txt 2>&1; tail -5 gpurun_out/r02_three_level_first.txt
timeout 600 python bench.py --sweep 4,6,8,10,12,16 --steps 50 --sweep-tag _base > gpurun_out/r02_sweep_base.log 2>&1; tail -3 gpurun_out/r02_sweep_base.log | cut -c1-600
python bench.py --sweep 12 --steps 3 --warmup 3 --sweep-tag _x > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:patch_kernel -s 4 -c 1 -o gpurun_out/r02_base_p12 python bench.py --sweep 12 --steps 3 --warmup 3 --sweep-tag _x > gpurun_out/ncu12.log 2>&1
python bench.py --sweep 16 --steps 3 --warmup 3 --sweep-tag _x > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:patch_kernel -s 4 -c 1 -o gpurun_out/r02_base_p16 python bench.py --sweep 16 --steps 3 --warmup 3 --sweep-tag _x > gpurun_out/ncu16.log 2>&1
ls -la gpurun_out | tail
