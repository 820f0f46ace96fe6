#!/bin/bash
# round 2, GPU visit 5: quaternion 8-product GEMM (qgemm.cu): parity, shapes, whole-solve A/B; DMMA + DADD co-issue probe
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/r02_05_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_05_pytest.log | cut -c1-1500
timeout 60 ./tools/fp64_peak | grep dadd | tee gpurun_out/r02_05_dmma_dadd.jsonl
timeout 200 python tools/gemm_probe.py 16384 2>&1 | tail -1 | tee gpurun_out/r02_05_gemm_probe.json | cut -c1-3000
timeout 200 python tools/probe_solve.py 16384 0 2 2>&1 | tail -1 | cut -c1-600 | tee gpurun_out/r02_05_probe.jsonl
ZQ_QGEMM=0 timeout 200 python tools/probe_solve.py 16384 0 1 2>&1 | tail -1 | cut -c1-600 | tee -a gpurun_out/r02_05_probe.jsonl
for c in 1 2 3; do ZQ_QGEMM=0 ZQ_3M_CFG=$c timeout 200 python tools/probe_solve.py 16384 0 1 2>&1 | tail -1 | cut -c1-600 | tee -a gpurun_out/r02_05_probe.jsonl; done
timeout 300 python bench.py --n2 8192 --steps 3 --warmup 3 --no-cpu 2>&1 | grep '^{' | cut -c1-2500 | tee gpurun_out/r02_05_bench8192.json
