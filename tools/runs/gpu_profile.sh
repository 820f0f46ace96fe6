#!/bin/bash
mkdir -p gpurun_out
timeout 1000 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_2048.csv python bench.py --n2 2048 --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_matvec|k_zgemm_mma" -c 10 -o gpurun_out/prof_r01 -f python tools/profile_kernels.py > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -5 gpurun_out/ncu_full.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cat gpurun_out/bench_default.json
