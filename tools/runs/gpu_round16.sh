#!/bin/bash
# round-1 session 3, call 1: parity after (a) re-gridded panel-dot CTAs in K1, (b) unrolled panel loops, (c) programmatic
# dependent launch of the per-column chain; A/B of ZQ_PDL on the mid-size, headline-size and batched workloads
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest(PDL=1) rc=$?"; tail -3 gpurun_out/pytest_gpu.log
ZQ_PDL=0 timeout 300 python -m pytest tests/test_gpu_zquatev.py tests/test_gpu_kernels.py -q -m gpu --timeout 300 -x > gpurun_out/pytest_gpu_pdl0.log 2>&1; echo "pytest(PDL=0) rc=$?"; tail -2 gpurun_out/pytest_gpu_pdl0.log
for pdl in 1 0; do
  ZQ_PDL=$pdl timeout 100 python tools/probe_solve.py 4096 0 2 2>&1 | tail -1 | cut -c1-420 | tee -a gpurun_out/probe16.jsonl
  ZQ_PDL=$pdl timeout 100 python tools/probe_solve.py 1024 0 2 2>&1 | tail -1 | cut -c1-420 | tee -a gpurun_out/probe16.jsonl
  ZQ_PDL=$pdl timeout 200 python tools/probe_solve.py 16384 0 1 2>&1 | tail -1 | cut -c1-420 | tee -a gpurun_out/probe16.jsonl
  ZQ_PDL=$pdl timeout 200 python tools/config45.py 5 256 2>&1 | tail -1 | cut -c1-700 | tee -a gpurun_out/probe16.jsonl
done
