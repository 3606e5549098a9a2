N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r02_bench_n${N}_e2e.json 2> gpurun_out/r02_bench_n${N}_e2e.err; echo "rc=$?"
tail -5 gpurun_out/r02_bench_n${N}_e2e.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_n${N}_e2e.json').read().strip().splitlines()[-1])
t=d['time_to_solution']
print(d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e'], t['seconds'], (d['parity'] or {}).get('peer_equals_nccl_bitwise'))"
