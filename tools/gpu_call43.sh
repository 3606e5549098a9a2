timeout 900 python bench.py > gpurun_out/r02_final_bench_n1.json 2> gpurun_out/r02_final_bench_n1.err
for n in 2 4; do
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 100 --warmup 10 > gpurun_out/r02_final_bench_n$n.json 2> gpurun_out/r02_final_bench_n$n.err
done
python - <<'PY'
import json
for n in (1,2,4):
    d=json.loads(open('gpurun_out/r02_final_bench_n%d.json'%n).read().strip().splitlines()[-1])
    t=d['time_to_solution']
    print(n, round(d['value'],2), round(d['ms_per_step'],4), round(d['roofline']['frac'],3), round(d['e2e']['value'],2), round(t['seconds'],4), t['outer_iterations'], t['inner_iterations'], t.get('setup_seconds'), d['clocks']['reasons'])
d=json.loads(open('gpurun_out/r02_final_bench_n1.json').read().strip().splitlines()[-1])
print(d['stokes']['time_to_solution']['seconds'], d['stokes']['time_to_solution']['gmres_iterations'], d['stokes'].get('cpu_baseline'))
PY
tail -2 gpurun_out/r02_final_bench_n1.err
