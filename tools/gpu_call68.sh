L=spectralelementmethod_b200/csrc/libsemk.so
cp $L /tmp/libsemk_cur.so
for rep in 1 2; do
  for v in cur pad; do
    if [ $v = pad ]; then cp tools/libsemk_pad.so $L; else cp /tmp/libsemk_cur.so $L; fi
    python bench.py --sweep 4,6,10 --steps 50 --warmup 5 --sweep-tag _c68_$v 2>&1 | grep "sweep p" | sed "s/^/$v /"
  done
done
cp /tmp/libsemk_cur.so $L
