#!/bin/bash
# round 2, GPU visit 6 (2 GPUs): per-panel column broadcast in the collective reduction (parity at world 2 + bench), handle /
# robustness tests, ncu warp-state capture of the quaternion GEMM
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/r02_06_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02_06_pytest.log | cut -c1-2500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 2 --no-e2e > gpurun_out/r02_06_bench2.json 2> gpurun_out/r02_06_bench2.err; echo "bench2 rc=$?"; grep '^{' gpurun_out/r02_06_bench2.json | cut -c1-2200; tail -5 gpurun_out/r02_06_bench2.err | cut -c1-1000
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_qgemm8|k_zgemm_3m" -c 4 -o gpurun_out/r02_06_qgemm -f python tools/profile_qgemm.py > gpurun_out/r02_06_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r02_06_ncu.log
