for mode in pair box column pair box; do
  python bench.py --sweep 16 --steps 50 --warmup 5 --ho-mode $mode --sweep-tag _c65_$mode 2>&1 | grep "sweep p" | sed "s/^/$mode /"
done
