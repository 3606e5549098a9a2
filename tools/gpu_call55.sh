# memcheck of the box gather on ragged tiles (empty slots) and of the batched host apply
compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_box.py -x -q -k "C-7-19-8-16 or C-3-25-6-8 or C-5-17-3-8 or batched" > gpurun_out/r02_c55_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -8 gpurun_out/r02_c55_memcheck.log
