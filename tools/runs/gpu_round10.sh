#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python tools/config45.py 5 256 2>&1 | tail -2 | tee gpurun_out/config5.json
timeout 300 python tools/config45.py 4 32768 2>&1 | tail -2 | tee gpurun_out/config4.json
