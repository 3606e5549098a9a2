timeout 900 python -m pytest tests/test_gpu_stokes.py -x -q > gpurun_out/r02_c31_pytest.log 2>&1; tail -6 gpurun_out/r02_c31_pytest.log
timeout 900 python bench.py > gpurun_out/r02_c31_bench.json 2> gpurun_out/r02_c31_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c31_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], d['roofline']['traffic'], d['e2e']['value'], d['time_to_solution']['seconds'], d['time_to_solution']['setup_seconds'])
c=d['condensed']; print({k:c[k] for k in ('host_numbering_seconds','setup_seconds','ms_per_apply')})
print(d['config']['setup_seconds'])
print(json.dumps(d['stokes'])[:1800])
PY
tail -3 gpurun_out/r02_c31_bench.err
