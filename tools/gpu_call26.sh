B="python bench.py --steps 3 --warmup 3 --no-condensed --pcg-iters 0 --cpu-sample 0 --e2e-steps 1"
$B > gpurun_out/r02_c26_plain.json 2> gpurun_out/r02_c26_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_apply.csv $B > gpurun_out/ncu_l.log 2>&1
cap() {  # name, kernel regex, skip, command...
  name=$1; k=$2; s=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -f -o /tmp/$name "$@" > gpurun_out/ncu_$name.log 2>&1
  python profiles/ncu_summary.py /tmp/$name.ncu-rep 14 > gpurun_out/r02_ncu_${name}_summary.txt 2>&1
  rm -f /tmp/$name.ncu-rep
}
cap apply_patch patch_kernel 4 $B
cap apply_shared shared_nodes_kernel 4 $B
S="python tests/stokes_bench.py 224 352 8 0 5"
$S > /dev/null 2>&1 && cap stokes_patch stokes_patch_kernel 3 $S
P="python bench.py --sweep 16 --steps 3 --warmup 3 --sweep-tag _x"
$P > /dev/null 2>&1 && cap pair_p16 ho_patch_kernel 4 $P
head -12 gpurun_out/r02_ncu_stokes_patch_summary.txt
