#!/bin/bash
# round 2, GPU visit 16 (8 GPUs): collective host-pointer pipeline at 8 ranks, 2n = 32768 (one device-resident solve, then the
# host-pointer entry with the pipeline on / off)
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29516 tools/dist_probe.py 16384 hostonly > gpurun_out/r02_16_probe_w8.jsonl 2> gpurun_out/r02_16_probe_w8.err; echo "probe rc=$?"
grep '^{' gpurun_out/r02_16_probe_w8.jsonl | cut -c1-700; tail -3 gpurun_out/r02_16_probe_w8.err | cut -c1-400
