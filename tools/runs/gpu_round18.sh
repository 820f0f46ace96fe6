#!/bin/bash
# round-1 session 3, call 3: K5 one-CTA reduction for n <= 256 (both reductions parametrised in the tests), threaded
# batched entry; config 5 at batch 1024 with K5 on/off
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-400
timeout 200 python tools/config45.py 5 1024 2>&1 | tail -1 | cut -c1-700 | tee -a gpurun_out/probe18.jsonl
ZQ_SMALL_N=0 timeout 200 python tools/config45.py 5 1024 2>&1 | tail -1 | cut -c1-700 | tee -a gpurun_out/probe18.jsonl
ZQ_BATCH_LANES=64 ZQ_BATCH_THREADS=12 timeout 200 python tools/config45.py 5 1024 2>&1 | tail -1 | cut -c1-700 | tee -a gpurun_out/probe18.jsonl
ZQ_BATCH_LANES=24 ZQ_BATCH_THREADS=4 timeout 200 python tools/config45.py 5 1024 2>&1 | tail -1 | cut -c1-700 | tee -a gpurun_out/probe18.jsonl
