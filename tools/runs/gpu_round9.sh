#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 520 --csv --log-file gpurun_out/launches_32768_window.csv python bench.py --n2 32768 --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/ncu_window.log 2>&1; echo "rc=$?"
python - <<'P'
import csv, collections
rows=list(csv.reader(open("gpurun_out/launches_32768_window.csv")))
hdr=None; agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    if "Kernel Name" in r: hdr=r; continue
    if hdr is None or len(r)<len(hdr): continue
    d=dict(zip(hdr,r))
    try: v=float(d["Metric Value"].replace(",",""))
    except: continue
    u=d.get("Metric Unit","")
    if u in("us","usecond"): v*=1e3
    elif u in ("ms","msecond"): v*=1e6
    k=d["Kernel Name"].split("(")[0].split("::")[-1][:40]
    agg[k][0]+=1; agg[k][1]+=v
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{k:40s} n={v[0]:4d} total_ms={v[1]*1e-6:9.3f} share={v[1]/tot*100:5.1f}% avg_us={v[1]/v[0]*1e-3:9.2f}")
P
