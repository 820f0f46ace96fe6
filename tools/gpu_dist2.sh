#!/bin/bash
# 2-GPU visit (tight timeouts: charged 2x)
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_dist.py -q -m gpu -x --timeout 200 > gpurun_out/pytest_dist.log 2>&1; echo "dist pytest rc=$?"; tail -4 gpurun_out/pytest_dist.log | cut -c1-600
for mode in 0 1; do
ZQ_DIST_NCCL=$mode timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2961$mode bench.py --gpus 2 --n2 8192 --steps 2 --warmup 2 --no-e2e --no-cpu > gpurun_out/bench_8192_g2_m$mode.json 2> gpurun_out/bench_8192_g2_m$mode.err; echo "nccl=$mode rc=$?"
python - <<P
import json
try:
    d=json.load(open("gpurun_out/bench_8192_g2_m$mode.json")); print(d["value"], d["phases_ms"], d["config"]["parallelism"][:140])
except Exception as e: print("no json", e)
P
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/bench_32768_g2.json 2> gpurun_out/bench_32768_g2.err; echo "rc=$?"
python - <<P
import json
try:
    d=json.load(open("gpurun_out/bench_32768_g2.json")); print(d["value"], d["phases_ms"], d["roofline"]["frac"])
except Exception as e: print("no json", e)
P
tail -3 gpurun_out/bench_32768_g2.err | cut -c1-300
