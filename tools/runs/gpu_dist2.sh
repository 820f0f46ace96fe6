#!/bin/bash
# 2-GPU visit (tight timeouts: charged 2x)
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_dist.py -q -m gpu -x --timeout 200 > gpurun_out/pytest_dist.log 2>&1; echo "dist pytest rc=$?"; tail -4 gpurun_out/pytest_dist.log | cut -c1-600
timeout 120 python tools/config45.py 5 256 2>&1 | tail -1 | tee gpurun_out/config5.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu > gpurun_out/bench_32768_g2.json 2> gpurun_out/bench_32768_g2.err; echo "rc=$?"
python - <<P
import json
for ln in open("gpurun_out/bench_32768_g2.json"):
    if ln.startswith("{"):
        d=json.loads(ln); print(d["value"], d["phases_ms"], d["roofline"]["frac"], d["roofline_fp64"]["frac"], d["e2e"]["value"], d["e2e"]["phases_ms"])
P
tail -3 gpurun_out/bench_32768_g2.err | cut -c1-300
