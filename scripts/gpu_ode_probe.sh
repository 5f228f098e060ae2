#!/bin/bash
mkdir -p gpurun_out
python scripts/ode_only.py; python scripts/ode_only.py --eager
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 160 --csv --log-file gpurun_out/launches_ode.csv python scripts/ode_only.py --eager > gpurun_out/ncu_ode.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/launches_ode.csv')))
hdr=None; agg=collections.OrderedDict(); seq=[]
for r in rows:
    if len(r)>5 and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); name=d['Kernel Name'][:60]
        try: v=float(d['Metric Value'].replace(',',''))
        except: continue
        a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v; seq.append((name[5:28], d['Grid Size'], round(v/1e3,1)))
for k,(n,t) in agg.items(): print(f"{n:4d} {t/1e3:10.1f} us total {t/n/1e3:8.2f} us/launch {k}")
print(seq[:14])
PY
