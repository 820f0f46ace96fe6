#!/bin/bash
# round 2, GPU visit 2: K1 with column-block runs (tpb sweep), handle / pipelined host mode, parity suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/r02_02_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_02_pytest.log | cut -c1-600
for t in 1 2 4; do ZQ_K1_TPB=$t timeout 120 python tools/k1_sweep.py 16384 20 2>&1 | tail -1 | tee -a gpurun_out/r02_02_k1sweep.jsonl | cut -c1-1500; done
timeout 200 python tools/probe_solve.py 16384 0 2 2>&1 | tail -1 | cut -c1-600 | tee gpurun_out/r02_02_probe.jsonl
ZQ_K1_TPB=1 timeout 200 python tools/probe_solve.py 16384 0 1 2>&1 | tail -1 | cut -c1-600 | tee -a gpurun_out/r02_02_probe.jsonl
ZQ_K1_TPB=2 timeout 200 python tools/probe_solve.py 16384 0 1 2>&1 | tail -1 | cut -c1-600 | tee -a gpurun_out/r02_02_probe.jsonl
timeout 900 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r02_02_bench.json 2> gpurun_out/r02_02_bench.err; echo "bench rc=$?"; grep '^{' gpurun_out/r02_02_bench.json | cut -c1-5000; tail -5 gpurun_out/r02_02_bench.err
