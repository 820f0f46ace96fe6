#!/bin/bash
# round 2, GPU visit 12 (2 GPUs): collective host-pointer pipeline (shared upload, sub-block exchange + download), split lower
# D&C levels, early panel push: parity worker (incl. the new host-pointer cases) + A/B probe at 2n = 32768
mkdir -p gpurun_out
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/dist_worker.py > gpurun_out/r02_12_dist.log 2>&1; echo "dist_worker rc=$?"
python - <<'PY'
import json
txt = open("gpurun_out/r02_12_dist.log").read()
i = txt.find("DIST_RESULT ")
if i < 0:
    print("no DIST_RESULT; tail:", txt[-3000:])
else:
    out = json.loads(txt[i + 12:].splitlines()[0])
    bad = [o for o in out if not o["ok"]]
    print("cases", len(out), "failed", len(bad))
    for o in bad: print("FAIL", json.dumps(o)[:600])
    for o in out:
        if "host-pointers" in o.get("mode", ""): print(json.dumps(o)[:400])
    if bad: print(txt[-1500:])
PY
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dist_probe.py 16384 > gpurun_out/r02_12_probe.jsonl 2> gpurun_out/r02_12_probe.err; echo "probe rc=$?"
grep '^{' gpurun_out/r02_12_probe.jsonl | cut -c1-700; tail -5 gpurun_out/r02_12_probe.err | cut -c1-400
