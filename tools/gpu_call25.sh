timeout 600 python -m pytest tests/test_gpu_multilevel.py -x -q -k "overlapped or halo_exchange" > gpurun_out/r02_c25_pytest.log 2>&1; tail -3 gpurun_out/r02_c25_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 10 --no-tts > gpurun_out/r02_c25_bench2.json 2> gpurun_out/r02_c25_bench2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c25_bench2.json').read().strip().splitlines()[-1])
print('N=2', d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['e2e']['value'], d['parity'] is not None)
PY
tail -3 gpurun_out/r02_c25_bench2.err
