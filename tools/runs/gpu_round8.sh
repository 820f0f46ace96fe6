#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_dc --csv --log-file gpurun_out/dc_launches.csv python tools/dc_probe.py 16384 > gpurun_out/dc_probe.log 2>&1; cat gpurun_out/dc_probe.log | tail -3
python - <<'P'
import csv, collections
rows=list(csv.reader(open("gpurun_out/dc_launches.csv")))
hdr=None; agg=collections.defaultdict(lambda:[0,0.0]); seq=[]
for r in rows:
    if "Kernel Name" in r: hdr=r; continue
    if hdr is None or len(r)<len(hdr): continue
    d=dict(zip(hdr,r))
    try: v=float(d["Metric Value"].replace(",",""))
    except: continue
    u=d.get("Metric Unit","")
    if u in("us","usecond"): v*=1e3
    elif u in ("ms","msecond"): v*=1e6
    k=d["Kernel Name"].split("(")[0].split("::")[-1]
    agg[k][0]+=1; agg[k][1]+=v; seq.append((k,v))
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{k:24s} n={v[0]:4d} total_ms={v[1]*1e-6:9.3f} share={v[1]/tot*100:5.1f}%")
print("total ms (2 solves)", tot*1e-6)
print("last level:", [(k, round(v*1e-6,3)) for k,v in seq[-10:]])
P
timeout 200 python tools/dc_probe.py 16384 2>&1 | tail -2
