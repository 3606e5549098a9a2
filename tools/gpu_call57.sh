C="python tools/condensed_apply.py 1024 6"
$C > gpurun_out/r02_c57_plain.log 2>&1 || exit 1
cat gpurun_out/r02_c57_plain.log
cap() {  # name, kernel regex, skip, command...
  name=$1; k=$2; s=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -f -o /tmp/$name "$@" > gpurun_out/ncu_$name.log 2>&1
  python profiles/ncu_summary.py /tmp/$name.ncu-rep 12 > gpurun_out/r02_ncu_${name}_summary.txt 2>&1
  rm -f /tmp/$name.ncu-rep
}
cap sc_matvec sc_matvec_kernel 3 $C
cap sc_node sc_node_kernel 3 $C
head -8 gpurun_out/r02_ncu_sc_matvec_summary.txt; head -8 gpurun_out/r02_ncu_sc_node_summary.txt
python -m pytest tests -x -q -m gpu > gpurun_out/r02_c57_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_c57_pytest.log
