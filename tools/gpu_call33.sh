timeout 1200 python -m pytest tests/test_gpu_multilevel.py tests/test_gpu_stokes.py -x -q > gpurun_out/r02_c33_pytest.log 2>&1; tail -6 gpurun_out/r02_c33_pytest.log
timeout 900 python bench.py > gpurun_out/r02_c33_bench.json 2> gpurun_out/r02_c33_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c33_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], d['time_to_solution']['seconds'])
print(d['condensed']['three_level_pcg'])
PY
tail -3 gpurun_out/r02_c33_bench.err
