# box kernel: correctness, then A/B against the table-driven kernel inside one box; PCIe probe
set -x
python -m pytest tests/test_gpu_box.py -x -q > gpurun_out/r02_c46_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_c46_pytest.log
tail -5 gpurun_out/r02_c46_pytest.log
for mode in column box column box; do
  python bench.py --sweep 8 --nx 1024 --steps 100 --warmup 10 --ho-mode $mode --sweep-tag _c46_${mode} 2>&1 | grep "sweep p"
done
python bench.py --sweep 6,10,12 --steps 50 --warmup 5 --ho-mode column --sweep-tag _c46_column_o 2>&1 | grep "sweep p"
python bench.py --sweep 6,10,12 --steps 50 --warmup 5 --ho-mode box --sweep-tag _c46_box_o 2>&1 | grep "sweep p"
python tools/pcie_probe.py > gpurun_out/r02_c46_pcie.log 2>&1; tail -12 gpurun_out/r02_c46_pcie.log
