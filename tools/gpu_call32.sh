timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/r02_c32_bench8.json 2> gpurun_out/r02_c32_bench8.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c32_bench8.json').read().strip().splitlines()[-1])
print('N=8', d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['e2e']['value'], d['time_to_solution'], d['parity'] is not None)
PY
tail -3 gpurun_out/r02_c32_bench8.err
