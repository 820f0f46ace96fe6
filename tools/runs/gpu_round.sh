#!/bin/bash
# One GPU visit: smoke, gpu tests, micro-benchmarks, quick bench.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
(nproc; lscpu | grep -E "Model name|Socket|Thread|Core") > gpurun_out/cpu.txt 2>&1
timeout 300 python __graft_entry__.py > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?" | tee -a gpurun_out/sanitizer.log; tail -5 gpurun_out/sanitizer.log
timeout 120 ./tools/fp64_peak > gpurun_out/fp64_peak.jsonl 2>&1
timeout 900 python tools/quick_bench.py > gpurun_out/quick_bench.jsonl 2>&1; echo "quick rc=$?"
tail -30 gpurun_out/quick_bench.jsonl
cat gpurun_out/fp64_peak.jsonl | tail -12
