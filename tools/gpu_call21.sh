timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "pair_kernel or 32_element or apply_vs_oracle_random" > gpurun_out/r02_c21_pytest.log 2>&1; tail -15 gpurun_out/r02_c21_pytest.log
for cfg in "10 8" "10 16" "12 4" "12 8" "16 4" "8 16" "8 8"; do set -- $cfg
  timeout 200 python bench.py --sweep $1 --pe $2 --ho-mode pair --steps 50 --sweep-tag _pair_p$1_pe$2 2>&1 >/dev/null | tail -1
done
for cfg in "4 32" "6 32" "4 16" "6 16" "3 32" "2 32"; do set -- $cfg
  timeout 200 python bench.py --sweep $1 --pe $2 --steps 50 --sweep-tag _col_p$1_pe$2 2>&1 >/dev/null | tail -1
done
timeout 200 python tests/stokes_bench.py 224 352 8 0 50 0 4 | cut -c1-700
timeout 200 python tests/stokes_bench.py 448 704 8 0 50 0 8 | cut -c1-700
