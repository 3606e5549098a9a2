timeout 900 python -m pytest tests/test_gpu_stokes.py tests/test_gpu_two_level.py -x -q > gpurun_out/r02_c24_pytest.log 2>&1; tail -8 gpurun_out/r02_c24_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py > gpurun_out/r02_c24_check.log 2>&1; tail -4 gpurun_out/r02_c24_check.log | cut -c1-1500
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r02_c24_bench2.json 2> gpurun_out/r02_c24_bench2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c24_bench2.json').read().strip().splitlines()[-1])
print('N=2', d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['e2e']['value'], d['time_to_solution']['seconds'], d['time_to_solution']['outer_iterations'], d['parity'] is not None)
PY
tail -5 gpurun_out/r02_c24_bench2.err
