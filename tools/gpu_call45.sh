SEMK_NO_GRAPH=1 python tests/ml_profile.py > gpurun_out/r02_c45_plain.log 2>&1 && SEMK_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 400 --csv --log-file gpurun_out/r02_c45_launches.csv python tests/ml_profile.py > gpurun_out/r02_c45_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02_c45_launches.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split('(')[0][-60:]
    t=float(r[-1])
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=t
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print("%-62s n=%4d  avg %8.1f us  share %5.1f %%"%(k,v[0],v[1]/v[0]/1e3,100*v[1]/tot))
PY
