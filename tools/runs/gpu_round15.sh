#!/bin/bash
# hoisted cp.async schedule in the GEMM kernels: parity + rates (compare profiles/r01_gemm_3m.jsonl) + full solve
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout 600 -x 2>&1 | tail -3
timeout 200 python tools/quick_bench.py gemm 2>&1 | grep gemm | tee gpurun_out/gemm_hoist.jsonl
timeout 200 python tools/probe_solve.py 16384 0 1 2>&1 | tail -1 | cut -c1-400 | tee -a gpurun_out/probe15.jsonl
