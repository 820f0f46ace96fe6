#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python tools/quick_bench.py 4096 2>&1 | grep -E '"n"|k1_n' | tee gpurun_out/k1_2dgrid.jsonl
timeout 300 bash tools/gpu_dropin.sh > /dev/null 2>&1; grep -E "===|Max dev|Errors|zheev|zquartev" gpurun_out/dropin.txt
timeout 400 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; python - <<'P'
import json
d=json.load(open("gpurun_out/bench_q.json"))
print(d["value"], d["phases_ms"], d["roofline"]["achieved"], d["roofline"]["k1_ms_per_step"])
P
