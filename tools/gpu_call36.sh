for cfg in "10 8" "10 4" "12 4" "12 16" "16 8"; do set -- $cfg
  timeout 200 python bench.py --sweep $1 --pe $2 --ho-mode column --steps 50 --sweep-tag _colx_p$1_pe$2 2>&1 >/dev/null | tail -1
done
