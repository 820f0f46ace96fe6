#!/bin/bash
# batched config 5: lanes x host-threads sweep
mkdir -p gpurun_out
nproc | tee gpurun_out/nproc.txt
for cfg in "64 12" "96 12" "96 16" "128 16" "128 24"; do
  set -- $cfg
  ZQ_BATCH_LANES=$1 ZQ_BATCH_THREADS=$2 timeout 200 python tools/config45.py 5 1024 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print(json.dumps({'lanes': $1, 'threads': $2, 'ms_per_matrix': d['ms_per_matrix'], 'seconds': d['seconds'], 'first': d['seconds_first_call'], 'ok': d['all_info_zero'], 'dev': d['max_eig_dev_rel']}))" | tee -a gpurun_out/probe21.jsonl
done
