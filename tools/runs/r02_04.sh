#!/bin/bash
# round 2, GPU visit 4 (8 GPUs): collective-solve parity at world = 8 and the 8-GPU bench line with the full-size quality block
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -q -m gpu --timeout 500 -k "8" > gpurun_out/r02_04_pytest8.log 2>&1; echo "pytest8 rc=$?"; tail -5 gpurun_out/r02_04_pytest8.log | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 2 --warmup 2 > gpurun_out/r02_04_bench8.json 2> gpurun_out/r02_04_bench8.err; echo "bench8 rc=$?"; grep '^{' gpurun_out/r02_04_bench8.json | cut -c1-6000; tail -5 gpurun_out/r02_04_bench8.err | cut -c1-1000
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 2 --warmup 2 --no-e2e > gpurun_out/r02_04_bench4.json 2> gpurun_out/r02_04_bench4.err; echo "bench4 rc=$?"; grep '^{' gpurun_out/r02_04_bench4.json | cut -c1-3000; tail -3 gpurun_out/r02_04_bench4.err | cut -c1-600
