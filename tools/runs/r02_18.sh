#!/bin/bash
# round 2, GPU visit 18 (1 GPU, the round's last seconds): ncu launch list (gpu__time_duration only, no replay) of ONE solve at 2n = 4096
mkdir -p gpurun_out
timeout 80 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_18_launches.csv python tools/ncu_one_solve.py 2048 > gpurun_out/r02_18_ncu.out 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/r02_18_ncu.out | cut -c1-300; wc -l gpurun_out/r02_18_launches.csv
