#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for cfg in 1 2 3; do echo "== GEMM cfg $cfg"; ZQ_GEMM_CFG=$cfg timeout 300 python tools/quick_bench.py gemm 2>&1 | tee gpurun_out/gemm_cfg$cfg.jsonl; done
timeout 300 python tools/quick_bench.py 1024 4096 2>&1 | head -3 | tee gpurun_out/quick_4096.jsonl
