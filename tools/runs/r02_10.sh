#!/bin/bash
# round 2, GPU visit 10: persistent quaternion GEMM (cross-tile operand pipeline) and BK 16 + staged C: parity, shapes, whole solve
mkdir -p gpurun_out
ZQ_Q8_PERSIST=1 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_zquatev.py tests/test_gpu_caller_side.py -q -m gpu --timeout 500 -k "qgemm8 or large_golden or congruence or properties_at_scale" > gpurun_out/r02_10_pytest_persist.log 2>&1; echo "pytest persist rc=$?"; tail -4 gpurun_out/r02_10_pytest_persist.log | cut -c1-1500
timeout 200 python tools/gemm_probe.py 16384 2>&1 | tail -1 >> gpurun_out/r02_10_gemm_probe.jsonl
ZQ_Q8_CFG=5 timeout 200 python tools/gemm_probe.py 16384 2>&1 | tail -1 >> gpurun_out/r02_10_gemm_probe.jsonl
ZQ_Q8_PERSIST=1 timeout 200 python tools/gemm_probe.py 16384 2>&1 | tail -1 >> gpurun_out/r02_10_gemm_probe.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/r02_10_gemm_probe.jsonl'):
    d=json.loads(l); print(d['env'], [(r['op'][:34], r['m'], r['ms'], r['canonical_tflops']) for r in d['gemm'] if 'q8' in r['op']])
PY
timeout 200 python tools/probe_solve.py 16384 0 1 2>&1 | tail -1 | cut -c1-600 | tee -a gpurun_out/r02_10_probe.jsonl
ZQ_Q8_PERSIST=1 timeout 200 python tools/probe_solve.py 16384 0 1 2>&1 | tail -1 | cut -c1-600 | tee -a gpurun_out/r02_10_probe.jsonl
ZQ_Q8_CFG=5 timeout 200 python tools/probe_solve.py 16384 0 1 2>&1 | tail -1 | cut -c1-600 | tee -a gpurun_out/r02_10_probe.jsonl
