for st in 8 16 32 64; do
python bench.py --steps 20 --warmup 3 --no-condensed --pcg-iters 0 --cpu-sample 0 --e2e-steps 10 --e2e-stages $st 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print($st, d['e2e']['value'], d['e2e']['matches_device_result'], d['value'])"
done
