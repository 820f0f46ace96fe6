#!/bin/bash
# A/B: K1 zigzag sweep (L2 reuse between consecutive columns) at 2n = 32768 and 2n = 8192
mkdir -p gpurun_out
for zz in 1 0; do
  ZQ_K1_ZIGZAG=$zz timeout 200 python tools/probe_solve.py 16384 64 1 2>&1 | tail -1 | cut -c1-400 | tee -a gpurun_out/probe14.jsonl
  ZQ_K1_ZIGZAG=$zz timeout 200 python tools/probe_solve.py 4096 0 3 2>&1 | tail -1 | cut -c1-400 | tee -a gpurun_out/probe14.jsonl
done
timeout 300 python -m pytest tests -q -m gpu --timeout 600 -x -k "tridiag or k1 or golden or reference_library" 2>&1 | tail -2
