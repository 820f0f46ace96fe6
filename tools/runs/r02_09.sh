#!/bin/bash
# round 2, GPU visit 9: full parity suite (property tests, caller-side products), smoke, compute-sanitizer (memcheck + racecheck) on the
# small code paths, drop-in latency table vs the reference, both bench arms as the driver runs them (shortened steps)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/r02_09_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_09_pytest.log | cut -c1-2000
timeout 300 python __graft_entry__.py > gpurun_out/r02_09_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_09_smoke.log | cut -c1-700
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_case.py > gpurun_out/r02_09_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02_09_memcheck.log | cut -c1-300
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_case.py > gpurun_out/r02_09_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r02_09_racecheck.log | cut -c1-300
timeout 600 python tools/dropin_latency.py 400 1000 2000 4096 2>&1 | grep '^{' | tee gpurun_out/r02_09_dropin.jsonl | cut -c1-400
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_09_bench.json 2> gpurun_out/r02_09_bench.err; echo "bench rc=$?"; grep '^{' gpurun_out/r02_09_bench.json | cut -c1-1500; tail -3 gpurun_out/r02_09_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_09_bench_ref.json 2> gpurun_out/r02_09_bench_ref.err; echo "ref rc=$?"; grep '^{' gpurun_out/r02_09_bench_ref.json | cut -c1-1500
