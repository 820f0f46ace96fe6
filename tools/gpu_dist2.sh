#!/bin/bash
# 2-GPU visit: parity of the collective solve, then the scaling pair N=1,2 at a mid size and at the headline size
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt; nvidia-smi topo -m >> gpurun_out/gpus.txt 2>&1; free -g >> gpurun_out/gpus.txt
timeout 900 python -m pytest tests/test_gpu_dist.py -q -m gpu -x --timeout 900 > gpurun_out/pytest_dist.log 2>&1; echo "dist pytest rc=$?"; tail -15 gpurun_out/pytest_dist.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --n2 8192 --steps 2 --warmup 2 --no-e2e --no-cpu > gpurun_out/bench_8192_g2.json 2> gpurun_out/bench_8192_g2.err; echo "rc=$?"; cat gpurun_out/bench_8192_g2.json | cut -c1-900; tail -3 gpurun_out/bench_8192_g2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu > gpurun_out/bench_32768_g2.json 2> gpurun_out/bench_32768_g2.err; echo "rc=$?"; cat gpurun_out/bench_32768_g2.json | cut -c1-1200; tail -3 gpurun_out/bench_32768_g2.err
timeout 300 python tools/quick_bench.py 99999 2>&1 | grep k1_n | tee gpurun_out/k1_2dgrid.jsonl
