#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/quality_report.py > gpurun_out/quality.jsonl 2>&1; cat gpurun_out/quality.jsonl
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 900 python tools/quick_bench.py 4096 > gpurun_out/quick_bench.jsonl 2>&1; echo "quick rc=$?"
cat gpurun_out/quick_bench.jsonl
