#!/bin/bash
mkdir -p gpurun_out
free -g > gpurun_out/mem.txt
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --n2 8192 --steps 2 --warmup 3 > gpurun_out/bench_8192.json 2> gpurun_out/bench_8192.err; echo "bench8192 rc=$?"
cat gpurun_out/bench_8192.json; tail -5 gpurun_out/bench_8192.err
timeout 1200 python bench.py --steps 1 --warmup 1 > gpurun_out/bench_32768.json 2> gpurun_out/bench_32768.err; echo "bench32768 rc=$?"
cat gpurun_out/bench_32768.json; tail -5 gpurun_out/bench_32768.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
cat gpurun_out/mem.txt
