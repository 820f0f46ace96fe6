#!/bin/bash
# round 2, GPU visit 13 (1 GPU): K1 with the L2 evict_first hint on the matrix stream (ZQ_K1_EVICT) -- whole-solve A/B at 2n = 32768
mkdir -p gpurun_out
timeout 200 python tools/probe_solve.py 16384 0 1 2>&1 | tail -1 | cut -c1-600 | tee -a gpurun_out/r02_13_probe.jsonl
ZQ_K1_EVICT=1 timeout 200 python tools/probe_solve.py 16384 0 1 2>&1 | tail -1 | cut -c1-600 | tee -a gpurun_out/r02_13_probe.jsonl
ZQ_K1_EVICT=6144 timeout 200 python tools/probe_solve.py 16384 0 1 2>&1 | tail -1 | cut -c1-600 | tee -a gpurun_out/r02_13_probe.jsonl
ZQ_K1_EVICT=1 timeout 200 python tools/probe_solve.py 4096 0 2 2>&1 | tail -1 | cut -c1-600 | tee -a gpurun_out/r02_13_probe.jsonl
timeout 200 python tools/probe_solve.py 4096 0 2 2>&1 | tail -1 | cut -c1-600 | tee -a gpurun_out/r02_13_probe.jsonl
