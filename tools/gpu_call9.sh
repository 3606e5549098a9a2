timeout 900 python -m pytest tests/test_gpu_condensed.py tests/test_gpu_multilevel.py tests/test_gpu_condensed_dense.py -x -q > gpurun_out/r02_c9_pytest.log 2>&1; tail -6 gpurun_out/r02_c9_pytest.log
timeout 600 python bench.py > gpurun_out/r02_c9_bench.json 2> gpurun_out/r02_c9_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c9_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['time_to_solution'])
c=d['condensed']; print({k:c[k] for k in ('host_numbering_seconds','setup_seconds','schur_pass_seconds','ms_per_apply')}, c['three_level_pcg'])
PY
tail -3 gpurun_out/r02_c9_bench.err
