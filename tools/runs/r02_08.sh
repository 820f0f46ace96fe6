#!/bin/bash
# round 2, GPU visit 8 (8 GPUs): parity at world = 8 with the per-panel broadcast + quaternion GEMMs, the 8-GPU bench line (both
# halves: device-resident and e2e), BASELINE config 4 (2n = 65536 values only) and config 5 (1024 x 2n = 512) over the 8 GPUs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -q -m gpu --timeout 500 -k "8" > gpurun_out/r02_08_pytest8.log 2>&1; echo "pytest8 rc=$?"; tail -5 gpurun_out/r02_08_pytest8.log | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 2 --warmup 2 > gpurun_out/r02_08_bench8.json 2> gpurun_out/r02_08_bench8.err; echo "bench8 rc=$?"; grep '^{' gpurun_out/r02_08_bench8.json | cut -c1-2600; tail -3 gpurun_out/r02_08_bench8.err | cut -c1-1000
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 tools/config4_dist.py 32768 2>&1 | grep '^{' | tee gpurun_out/r02_08_config4_8gpu.json | cut -c1-1500
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 tools/config5_dist.py 1024 2>&1 | grep '^{' | tee gpurun_out/r02_08_config5_8gpu.json | cut -c1-1000
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 2 --warmup 2 --no-e2e > gpurun_out/r02_08_bench4.json 2> gpurun_out/r02_08_bench4.err; echo "bench4 rc=$?"; grep '^{' gpurun_out/r02_08_bench4.json | cut -c1-1500
