# final box build: PCIe probe 2, ncu launch list + full captures of the two apply kernels, default bench line
set -x
python tools/pcie_probe2.py > gpurun_out/r02_c50_pcie2.log 2>&1; cat gpurun_out/r02_c50_pcie2.log
B3="python bench.py --steps 3 --warmup 3 --no-condensed --pcg-iters 0 --cpu-sample 0 --e2e-steps 1 --no-tts"
$B3 > gpurun_out/r02_c50_plain.json 2> gpurun_out/r02_c50_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_apply_box.csv $B3 > gpurun_out/ncu_l.log 2>&1
cap() {  # name, kernel regex, skip, command...
  name=$1; k=$2; s=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -f -o /tmp/$name "$@" > gpurun_out/ncu_$name.log 2>&1
  python profiles/ncu_summary.py /tmp/$name.ncu-rep 14 > gpurun_out/r02_ncu_${name}_summary.txt 2>&1
  rm -f /tmp/$name.ncu-rep
}
cap box_patch patch_kernel 4 $B3
cap box_shared shared_nodes_kernel 4 $B3
head -12 gpurun_out/r02_ncu_box_patch_summary.txt
python bench.py > gpurun_out/r02_bench_n1_box.json 2> gpurun_out/r02_bench_n1_box.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_n1_box.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel'], d['roofline']['traffic'], d['e2e'], d['time_to_solution']['seconds'])"
