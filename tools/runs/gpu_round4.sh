#!/bin/bash
mkdir -p gpurun_out
for nb in 32 48 64; do
timeout 600 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --nb $nb > gpurun_out/bench_nb$nb.json 2> gpurun_out/bench_nb$nb.err; echo "nb=$nb rc=$?"
python - <<P
import json
d=json.load(open("gpurun_out/bench_nb$nb.json"))
print($nb, d["value"], d["phases_ms"], d["roofline"]["achieved"], d["roofline"]["k1_ms_per_step"])
P
done
# launch list (cold-cache, serialised): shares per kernel at 2n=8192
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_8192.csv python bench.py --n2 8192 --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches rc=$?"
python - <<'P'
import csv, collections
rows=list(csv.reader(open("gpurun_out/launches_8192.csv")))
hdr=None; agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    if "Kernel Name" in r: hdr=r; continue
    if hdr is None or len(r)<len(hdr): continue
    d=dict(zip(hdr,r))
    try: v=float(d["Metric Value"].replace(",",""))
    except: continue
    unit=d.get("Metric Unit","")
    if unit=="us": v*=1e3
    elif unit=="ms": v*=1e6
    elif unit=="s": v*=1e9
    k=d["Kernel Name"][:60]
    agg[k][0]+=1; agg[k][1]+=v
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print(f"{k:60s} n={v[0]:6d} total_ms={v[1]*1e-6:10.3f} share={v[1]/tot*100:5.1f}% avg_us={v[1]/v[0]*1e-3:8.2f}")
P
