#!/bin/bash
# K5 alone: phase split of single device-resident solves at n = 256 / 128 with the one-CTA reduction and the chain
mkdir -p gpurun_out
for sn in 256 0; do
  ZQ_SMALL_N=$sn timeout 100 python tools/probe_solve.py 256 0 3 2>&1 | tail -1 | cut -c1-420 | tee -a gpurun_out/probe19.jsonl
  ZQ_SMALL_N=$sn timeout 100 python tools/probe_solve.py 128 0 2 2>&1 | tail -1 | cut -c1-420 | tee -a gpurun_out/probe19.jsonl
done
