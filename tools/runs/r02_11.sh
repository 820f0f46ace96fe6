#!/bin/bash
# round 2, GPU visit 11 (2 GPUs): two-panel back-transformation on the quaternion GEMM + persistent GEMM as defaults: full parity suite
# (incl. the 2-rank collective solve), whole-solve A/B, 1- and 2-GPU bench lines
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/r02_11_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_11_pytest.log | cut -c1-2000
timeout 200 python tools/probe_solve.py 16384 0 1 2>&1 | tail -1 | cut -c1-600 | tee -a gpurun_out/r02_11_probe.jsonl
ZQ_BT_PAIR=0 timeout 200 python tools/probe_solve.py 16384 0 1 2>&1 | tail -1 | cut -c1-600 | tee -a gpurun_out/r02_11_probe.jsonl
timeout 200 python tools/probe_solve.py 16384 2048 1 2>&1 | tail -1 | cut -c1-600 | tee -a gpurun_out/r02_11_probe.jsonl
ZQ_BT_PAIR=0 timeout 200 python tools/probe_solve.py 16384 2048 1 2>&1 | tail -1 | cut -c1-600 | tee -a gpurun_out/r02_11_probe.jsonl
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 2 > gpurun_out/r02_11_bench2.json 2> gpurun_out/r02_11_bench2.err; echo "bench2 rc=$?"; grep '^{' gpurun_out/r02_11_bench2.json | cut -c1-1800; tail -3 gpurun_out/r02_11_bench2.err | cut -c1-600
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_11_bench1.json 2> gpurun_out/r02_11_bench1.err; echo "bench1 rc=$?"; grep '^{' gpurun_out/r02_11_bench1.json | cut -c1-1800
timeout 100 python bench.py --n2 8192 --steps 3 --warmup 3 --no-cpu 2>/dev/null | grep '^{' | cut -c1-1500 | tee gpurun_out/r02_11_bench8192.json
