set -x
python -m pytest tests/test_gpu_box.py -x -q > gpurun_out/r02_c58_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_c58_pytest.log
for mode in column box column box; do
  python bench.py --sweep 2,3,4 --steps 100 --warmup 10 --ho-mode $mode --sweep-tag _c58_$mode 2>&1 | grep "sweep p"
done
B3="python bench.py --steps 3 --warmup 3 --no-condensed --pcg-iters 0 --cpu-sample 0 --e2e-steps 1 --no-tts"
cap() {  # name, kernel regex, skip, command...
  name=$1; k=$2; s=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -f -o /tmp/$name "$@" > gpurun_out/ncu_$name.log 2>&1
  python profiles/ncu_summary.py /tmp/$name.ncu-rep 14 > gpurun_out/r02_ncu_${name}_summary.txt 2>&1
  rm -f /tmp/$name.ncu-rep
}
cap box_patch patch_kernel 4 $B3
cap box_shared shared_nodes_kernel 4 $B3
head -4 gpurun_out/r02_ncu_box_patch_summary.txt
