B3="python bench.py --steps 3 --warmup 3 --no-condensed --pcg-iters 0 --cpu-sample 0 --e2e-steps 1 --no-tts"
ncu --set full --clock-control none --import-source on -k regex:patch_kernel -s 4 -c 1 -f -o /tmp/boxp $B3 > gpurun_out/ncu_boxp.log 2>&1
python tools/ncu_smem_lines.py /tmp/boxp.ncu-rep 60 > gpurun_out/r02_ncu_box_patch_smem_lines.txt 2>&1
head -80 gpurun_out/r02_ncu_box_patch_smem_lines.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
