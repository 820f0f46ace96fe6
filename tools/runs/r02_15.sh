#!/bin/bash
# round 2, GPU visit 15 (1 GPU): validation of the final library of this session -- full parity suite, smoke, bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 --durations=8 > gpurun_out/r02_15_pytest.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/r02_15_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_15_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_15_smoke.txt | cut -c1-600
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_15_bench1.json 2> gpurun_out/r02_15_bench1.err; echo "bench1 rc=$?"; grep '^{' gpurun_out/r02_15_bench1.json | cut -c1-2500; tail -3 gpurun_out/r02_15_bench1.err | cut -c1-400
timeout 100 python bench.py --n2 8192 --steps 3 --warmup 3 --no-cpu 2>/dev/null | grep '^{' | cut -c1-900 | tee gpurun_out/r02_15_bench8192.json
