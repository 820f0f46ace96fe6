#!/bin/bash
# round 2, GPU visit 3 (2 GPUs): parity suite incl. the 2-rank collective solve (peer exchange fused into the 3-launch chain),
# K1 runs of 8 column blocks, host-mode pipeline check, 2-GPU bench with the full-size quality block
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/r02_03_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_03_pytest.log | cut -c1-1500
ZQ_K1_TPB=8 timeout 120 python tools/k1_sweep.py 16384 20 2>&1 | tail -1 | tee -a gpurun_out/r02_03_k1sweep.jsonl | cut -c1-1500
timeout 120 python tools/k1_sweep.py 16384 20 2>&1 | tail -1 | tee -a gpurun_out/r02_03_k1sweep.jsonl | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 2 > gpurun_out/r02_03_bench2.json 2> gpurun_out/r02_03_bench2.err; echo "bench2 rc=$?"; grep '^{' gpurun_out/r02_03_bench2.json | cut -c1-5000; tail -5 gpurun_out/r02_03_bench2.err | cut -c1-1000
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r02_03_bench1.json 2> gpurun_out/r02_03_bench1.err; echo "bench1 rc=$?"; grep '^{' gpurun_out/r02_03_bench1.json | cut -c1-6000; tail -5 gpurun_out/r02_03_bench1.err
