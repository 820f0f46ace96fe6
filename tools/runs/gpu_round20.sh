#!/bin/bash
# K5 v2 (all loads of a column issued before the arithmetic): parity, single-solve phases, batch of 1024
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-400
timeout 100 python tools/probe_solve.py 256 0 3 2>&1 | tail -1 | cut -c1-420 | tee -a gpurun_out/probe20.jsonl
timeout 100 python tools/probe_solve.py 128 0 2 2>&1 | tail -1 | cut -c1-420 | tee -a gpurun_out/probe20.jsonl
timeout 200 python tools/config45.py 5 1024 2>&1 | tail -1 | cut -c1-700 | tee -a gpurun_out/probe20.jsonl
ZQ_BATCH_LANES=64 ZQ_BATCH_THREADS=12 timeout 200 python tools/config45.py 5 1024 2>&1 | tail -1 | cut -c1-700 | tee -a gpurun_out/probe20.jsonl
