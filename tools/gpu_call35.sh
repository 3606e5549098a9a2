timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c35_smoke.log 2>&1; tail -4 gpurun_out/r02_c35_smoke.log
timeout 1700 python -m pytest tests -x -q -m gpu > gpurun_out/r02_c35_pytest.log 2>&1; tail -5 gpurun_out/r02_c35_pytest.log
