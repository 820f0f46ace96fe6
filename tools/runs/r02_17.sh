#!/bin/bash
# round 2, GPU visit 17 (1 GPU, last of the round): (1) the full parity suite on the default path after the refactor (combos in the
# shared header, solver wiring); (2) the same suite with ZQ_Q8X=1 (pre-combined-operand GEMM for K4 and the K6 update);
# (3) whole-solve A/B at 2n = 32768; (4) a bench line with ZQ_Q8X=1 if it wins
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout 500 -x > gpurun_out/r02_17_pytest_default.log 2>&1; echo "pytest default rc=$?"; tail -4 gpurun_out/r02_17_pytest_default.log | cut -c1-400
ZQ_Q8X=1 timeout 600 python -m pytest tests -q -m gpu --timeout 500 > gpurun_out/r02_17_pytest_q8x.log 2>&1; echo "pytest q8x rc=$?"; tail -12 gpurun_out/r02_17_pytest_q8x.log | cut -c1-600
timeout 100 python tools/probe_solve.py 16384 0 1 2>&1 | tail -1 | cut -c1-600 | tee -a gpurun_out/r02_17_probe.jsonl
ZQ_Q8X=1 timeout 100 python tools/probe_solve.py 16384 0 1 2>&1 | tail -1 | cut -c1-600 | tee -a gpurun_out/r02_17_probe.jsonl
WIN=$(python - <<'PY'
import json
r = [json.loads(l) for l in open("gpurun_out/r02_17_probe.jsonl") if l.startswith("{")]
try:
    a, b = r[-2]["phases_ms"]["device_total"], r[-1]["phases_ms"]["device_total"]
    print(1 if (r[-1]["info"] == 0 and b < 0.995 * a) else 0)
except Exception:
    print(0)
PY
)
echo "q8x wins: $WIN"
if [ "$WIN" = "1" ]; then
  ZQ_Q8X=1 timeout 400 python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/r02_17_bench_q8x.json 2> gpurun_out/r02_17_bench_q8x.err; echo "bench q8x rc=$?"; grep '^{' gpurun_out/r02_17_bench_q8x.json | cut -c1-1200; tail -2 gpurun_out/r02_17_bench_q8x.err | cut -c1-300
fi
