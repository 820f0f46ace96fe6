#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python tools/quick_bench.py gemm 2>&1 | tee gpurun_out/gemm_3m.jsonl
timeout 300 python tools/quality_report.py 2>&1 | tail -6 | cut -c1-300
timeout 400 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; python - <<'P'
import json
for ln in open("gpurun_out/bench_q.json"):
    if ln.startswith("{"):
        d=json.loads(ln); print(d["value"], d["phases_ms"], d["roofline"]["frac"], d["roofline_fp64"]["achieved"])
P
