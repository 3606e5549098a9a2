python -m pytest tests/test_gpu_parity.py -x -q -k "pair or random" > gpurun_out/r02_c67_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_c67_pytest.log
python bench.py --sweep 16 --steps 50 --warmup 5 --sweep-tag _c67 2>&1 | grep "sweep p"
python bench.py --sweep 16 --steps 50 --warmup 5 --sweep-tag _c67 2>&1 | grep "sweep p"
