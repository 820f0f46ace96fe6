#!/bin/bash
# round 2, GPU visit 14 (4 GPUs): the round's multi-GPU changes at 4 ranks -- parity worker (world 4) and the short A/B probe at
# 2n = 32768 (three sub-blocks per rank, D&C levels of 1 / 2 / 4 blocks split, several senders to rank 0)
mkdir -p gpurun_out
W=${1:-4}
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29514 tests/dist_worker.py > gpurun_out/r02_14_dist_w$W.log 2>&1; echo "dist_worker rc=$?"
python - $W <<'PY'
import json, sys
W = sys.argv[1]
txt = open(f"gpurun_out/r02_14_dist_w{W}.log").read()
i = txt.find("DIST_RESULT ")
if i < 0:
    print("no DIST_RESULT; tail:", txt[-3000:])
else:
    out = json.loads(txt[i + 12:].splitlines()[0])
    bad = [o for o in out if not o["ok"]]
    print("cases", len(out), "failed", len(bad))
    for o in bad: print("FAIL", json.dumps(o)[:600])
    if bad: print(txt[-1500:])
PY
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29515 tools/dist_probe.py 16384 short > gpurun_out/r02_14_probe_w$W.jsonl 2> gpurun_out/r02_14_probe_w$W.err; echo "probe rc=$?"
grep '^{' gpurun_out/r02_14_probe_w$W.jsonl | cut -c1-700; tail -3 gpurun_out/r02_14_probe_w$W.err | cut -c1-400
