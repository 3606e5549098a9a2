# box kernel as the automatic choice: full GPU suite, config-2 A/B in one box, ncu captures
set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r02_c48_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02_c48_pytest.log
B="python bench.py --steps 100 --warmup 10 --no-condensed --pcg-iters 0 --cpu-sample 0 --e2e-steps 2 --no-tts"
for mode in column box column box; do
  SEMK_APPLY_MODE=$mode $B 2> gpurun_out/r02_c48_ab_$mode.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$mode', d['roofline']['kernel'], d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'])"
done
B3="python bench.py --steps 3 --warmup 3 --no-condensed --pcg-iters 0 --cpu-sample 0 --e2e-steps 1 --no-tts"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_apply_box.csv $B3 > gpurun_out/ncu_l.log 2>&1
cap() {  # name, kernel regex, skip, command...
  name=$1; k=$2; s=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -f -o /tmp/$name "$@" > gpurun_out/ncu_$name.log 2>&1
  python profiles/ncu_summary.py /tmp/$name.ncu-rep 14 > gpurun_out/r02_ncu_${name}_summary.txt 2>&1
  rm -f /tmp/$name.ncu-rep
}
cap box_patch patch_kernel 4 $B3
cap box_shared shared_nodes_kernel 4 $B3
head -30 gpurun_out/r02_ncu_box_patch_summary.txt
