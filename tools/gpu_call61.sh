# final verification: full GPU suite, smoke, default bench line
python -m pytest tests -x -q -m gpu > gpurun_out/r02_c61_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_c61_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c61_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02_c61_smoke.log
python bench.py > gpurun_out/r02_bench_n1_last.json 2> gpurun_out/r02_bench_n1_last.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_n1_last.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'], d['e2e']['value'], d['e2e']['single_call_value'], d['time_to_solution']['seconds'], d['condensed']['traffic'], d['stokes']['time_to_solution']['seconds'])"
