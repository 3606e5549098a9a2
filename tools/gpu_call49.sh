# box gather with / without the index-block copy (two libraries), config-2 A/B; batched e2e
set -x
python -m pytest tests/test_gpu_box.py -x -q > gpurun_out/r02_c49_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02_c49_pytest.log
L=spectralelementmethod_b200/csrc/libsemk.so
cp $L /tmp/libsemk_noskip.so
B="python bench.py --steps 200 --warmup 20 --no-condensed --pcg-iters 0 --cpu-sample 0 --e2e-steps 8 --no-tts"
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', d['roofline']['kernel'][:34], 'ms/step %.4f value %.2f frac %.4f e2e %.3f single %s' % (d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e'].get('single_call_value')))"; }
for rep in 1 2; do
  cp /tmp/libsemk_noskip.so $L
  SEMK_APPLY_MODE=column $B 2> gpurun_out/r02_c49_col.err | show column
  SEMK_APPLY_MODE=box $B 2> gpurun_out/r02_c49_box.err | show box_noskip
  cp tools/libsemk_skip.so $L
  SEMK_APPLY_MODE=box $B 2> gpurun_out/r02_c49_boxskip.err | show box_skip
done
cp /tmp/libsemk_noskip.so $L
