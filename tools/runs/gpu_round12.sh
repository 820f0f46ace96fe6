#!/bin/bash
mkdir -p gpurun_out
for cfg in "0 64" "1 64" "1 32"; do
set -- $cfg
ZQ_L2_PERSIST=$1 timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --nb $2 > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err
python - <<P
import json
for ln in open("gpurun_out/bench_ab.json"):
    if ln.startswith("{"):
        d=json.loads(ln); print("L2=$1 nb=$2", round(d["value"],3), {k:round(v,1) for k,v in d["phases_ms"].items() if k in("tridiag","tridiag_eig","backtransform")}, round(d["roofline"]["frac"],3), round(d["roofline"]["k1_ms_per_step"],1))
P
done
ZQ_L2_PERSIST=1 timeout 100 python tools/quick_bench.py 4096 2>&1 | head -1 | cut -c1-400
ZQ_L2_PERSIST=0 timeout 100 python tools/quick_bench.py 4096 2>&1 | head -1 | cut -c1-400
