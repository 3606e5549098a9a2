# persistent interface kernel with descriptor prefetch vs the one-shot kernel: tests, A/B, launch list
set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_box.py -x -q > gpurun_out/r02_c54_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02_c54_pytest.log
L=spectralelementmethod_b200/csrc/libsemk.so
cp $L /tmp/libsemk_new.so
B="python bench.py --steps 200 --warmup 20 --no-condensed --pcg-iters 0 --cpu-sample 0 --e2e-steps 2 --no-tts"
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', 'ms/step %.4f value %.2f frac %.4f' % (d['ms_per_step'], d['value'], d['roofline']['frac']))"; }
for rep in 1 2; do
  cp tools/libsemk_oldiface.so $L; $B 2> /dev/null | show old_iface
  cp /tmp/libsemk_new.so $L;       $B 2> /dev/null | show new_iface
done
B3="python bench.py --steps 3 --warmup 3 --no-condensed --pcg-iters 0 --cpu-sample 0 --e2e-steps 1 --no-tts"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_c54_launches.csv $B3 > gpurun_out/ncu_l.log 2>&1
grep "shared_nodes_kernel" gpurun_out/r02_c54_launches.csv | tail -4
