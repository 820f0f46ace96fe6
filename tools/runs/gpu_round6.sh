#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python tools/quick_bench.py gemm 2>&1 | tee gpurun_out/gemm_auto.jsonl
for nb in 32 64; do
timeout 600 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --nb $nb > gpurun_out/bench_nb$nb.json 2> gpurun_out/bench_nb$nb.err; echo "nb=$nb rc=$?"
python - <<P
import json
d=json.load(open("gpurun_out/bench_nb$nb.json"))
print($nb, d["value"], d["phases_ms"], d["roofline"]["achieved"], d["roofline"]["k1_ms_per_step"])
P
done
