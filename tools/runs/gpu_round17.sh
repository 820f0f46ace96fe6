#!/bin/bash
# 2-GPU visit (charged 2x): parity of the collective solve with programmatic dependent launch in the peer-exchange
# chain, then ZQ_PDL A/B at 2n=16384 and the headline size once
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_dist.py -q -m gpu -x --timeout 200 > gpurun_out/pytest_dist.log 2>&1; echo "dist pytest rc=$?"; tail -4 gpurun_out/pytest_dist.log | cut -c1-600
run() {  # $1 = ZQ_PDL, $2 = n2, $3 = out
  ZQ_PDL=$1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 2 --n2 $2 --steps 1 --warmup 1 --no-cpu --no-e2e > $3 2> $3.err; echo "rc=$?"
  python - "$3" <<P
import json, sys
for ln in open(sys.argv[1]):
    if ln.startswith("{"):
        d=json.loads(ln); print(d["config"]["n2"], d["value"], d["phases_ms"], d["roofline"]["frac"])
P
  tail -2 $3.err | cut -c1-300
}
run 1 16384 gpurun_out/bench_16384_g2_pdl1.json
run 0 16384 gpurun_out/bench_16384_g2_pdl0.json
run 1 32768 gpurun_out/bench_32768_g2.json
