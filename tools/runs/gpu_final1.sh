#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_matvec|k_zgemm_3m" -c 10 -o gpurun_out/prof_r01b -f python tools/profile_kernels.py > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_1024.csv python bench.py --n2 1024 --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 python bench.py > gpurun_out/bench_final_1gpu.json 2> gpurun_out/bench_final_1gpu.err; echo "bench rc=$?"; grep '^{' gpurun_out/bench_final_1gpu.json | cut -c1-2600
timeout 200 python bench.py --impl reference > gpurun_out/bench_final_ref.json 2>&1; grep '^{' gpurun_out/bench_final_ref.json | cut -c1-600
timeout 100 python bench.py --n2 8192 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_final_8192.json 2> /dev/null; grep '^{' gpurun_out/bench_final_8192.json | cut -c1-1200
