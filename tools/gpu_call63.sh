# is the time to solution sensitive to the threaded plan builder earlier in the process?
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
t=d['time_to_solution']
print('$1', 'setup %.2f' % d['config']['setup_seconds'], 'tts %.4f' % t['seconds'], t['outer_iterations'], t['inner_iterations'], 'value %.1f' % d['value'])"; }
B="python bench.py --steps 50 --warmup 5 --pcg-iters 0 --cpu-sample 0 --e2e-steps 2"
nproc
$B 2>/dev/null | show threads_all
SEMK_HOST_THREADS=1 $B 2>/dev/null | show threads_1
$B 2>/dev/null | show threads_all
