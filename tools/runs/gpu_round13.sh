#!/bin/bash
# round-1 session 2, call 1: parity of the split-K back-transformation + graph-replayed batches, smoke, A/B probes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log | cut -c1-400
for sk in 1 0; do
  ZQ_BT_SPLITK=$sk timeout 200 python tools/probe_solve.py 16384 0 1 2>&1 | tail -1 | cut -c1-500 | tee -a gpurun_out/probe13.jsonl
  ZQ_BT_SPLITK=$sk timeout 200 python tools/probe_solve.py 16384 2048 1 2>&1 | tail -1 | cut -c1-500 | tee -a gpurun_out/probe13.jsonl
done
for g in 1 0; do
  ZQ_BATCH_GRAPH=$g timeout 200 python tools/config45.py 5 256 2>&1 | tail -1 | cut -c1-700 | tee -a gpurun_out/probe13.jsonl
done
