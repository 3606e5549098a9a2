python bench.py --sweep 2,3,4,6,8,10,12,16 --steps 50 --warmup 5 --sweep-tag _box 2> gpurun_out/r02_sweep_box.err | tail -1 | cut -c1-200
grep "sweep p" gpurun_out/r02_sweep_box.err
