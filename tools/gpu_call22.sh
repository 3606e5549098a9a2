timeout 1700 python -m pytest tests -x -q -m gpu > gpurun_out/r02_c22_pytest.log 2>&1; tail -8 gpurun_out/r02_c22_pytest.log
timeout 600 python bench.py > gpurun_out/r02_c22_bench.json 2> gpurun_out/r02_c22_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c22_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['time_to_solution'])
print(d['stokes'])
PY
tail -3 gpurun_out/r02_c22_bench.err
