python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r02_bench_n2_box.json 2> gpurun_out/r02_bench_n2_box.err; echo "rc=$?"
tail -3 gpurun_out/r02_bench_n2_box.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_n2_box.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel'], d['e2e']['value'], d['time_to_solution'], d['parity'])"
