# does the 16-byte alignment of the by-value D matrix (kernel parameter) matter?  current library: D at
# offset 224 (aligned); libsemk_pad.so: one more 8-byte parameter in front of it (offset 232)
L=spectralelementmethod_b200/csrc/libsemk.so
cp $L /tmp/libsemk_cur.so
for rep in 1 2; do
  for v in cur pad; do
    if [ $v = pad ]; then cp tools/libsemk_pad.so $L; else cp /tmp/libsemk_cur.so $L; fi
    python bench.py --sweep 8,12,16 --steps 50 --warmup 5 --sweep-tag _c66_$v 2>&1 | grep "sweep p" | sed "s/^/$v /"
  done
done
cp /tmp/libsemk_cur.so $L
