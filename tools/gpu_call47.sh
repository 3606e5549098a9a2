# box kernel variants (SEMK_BOX_BITS: 1 cp.async gather, 2 arithmetic write-out, 4 arithmetic LDG gather)
set -x
for bits in 1 2 4 3 6; do
  SEMK_BOX_BITS=$bits python -m pytest tests/test_gpu_box.py -x -q -k "S-8-24-8-16 or C-7-19-8-16 or S-4-16-12-8 or C-6-17-8-16" > gpurun_out/r02_c47_pytest_$bits.log 2>&1; echo "bits $bits pytest rc=$?"
  tail -3 gpurun_out/r02_c47_pytest_$bits.log
done
python bench.py --sweep 8 --nx 1024 --steps 100 --warmup 10 --ho-mode column --sweep-tag _c47_column 2>&1 | grep "sweep p"
for bits in 1 2 4 3 6 1 4; do
  SEMK_BOX_BITS=$bits python bench.py --sweep 8 --nx 1024 --steps 100 --warmup 10 --ho-mode box --sweep-tag _c47_box$bits 2>&1 | grep "sweep p"
done
python bench.py --sweep 8 --nx 1024 --steps 100 --warmup 10 --ho-mode column --sweep-tag _c47_column 2>&1 | grep "sweep p"
for bits in 1 4 2; do
  SEMK_BOX_BITS=$bits python bench.py --sweep 6,10,12 --steps 50 --warmup 5 --ho-mode box --sweep-tag _c47_box${bits}_o 2>&1 | grep "sweep p"
done
