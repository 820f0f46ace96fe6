#!/bin/bash
# 8-GPU visit (charged 8x: keep it short)
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/gpus8.txt; free -g >> gpurun_out/gpus8.txt; nproc >> gpurun_out/gpus8.txt
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29708 tests/dist_worker.py > gpurun_out/dist8_worker.log 2>&1; echo "dist8 worker rc=$?"; grep DIST_RESULT gpurun_out/dist8_worker.log | cut -c1-1500
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29718 bench.py --gpus 8 --steps 1 --warmup 1 --no-cpu > gpurun_out/bench_32768_g8.json 2> gpurun_out/bench_32768_g8.err; echo "bench8 rc=$?"
grep '^{' gpurun_out/bench_32768_g8.json | cut -c1-3000
tail -3 gpurun_out/bench_32768_g8.err | cut -c1-300
cat gpurun_out/gpus8.txt
