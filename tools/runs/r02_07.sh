#!/bin/bash
# round 2, GPU visit 7: quaternion GEMM pipeline variants (A/B on shapes and whole solve), config 4 on one GPU (eigenvalues saved for
# the 1-vs-8 comparison), ncu --set full samples of K1 over a ladder of trailing sizes (DRAM traffic vs algorithmic bytes)
mkdir -p gpurun_out
for c in 0 1 2 3; do ZQ_Q8_CFG=$c timeout 200 python tools/gemm_probe.py 16384 2>&1 | tail -1 >> gpurun_out/r02_07_gemm_probe.jsonl; done; cut -c1-1200 gpurun_out/r02_07_gemm_probe.jsonl
for c in 1 3; do ZQ_Q8_CFG=$c timeout 200 python tools/probe_solve.py 16384 0 1 2>&1 | tail -1 | cut -c1-600 | tee -a gpurun_out/r02_07_probe.jsonl; done
timeout 600 python tools/config4_dist.py 32768 --save 2>&1 | tail -1 | tee gpurun_out/r02_07_config4_1gpu.json | cut -c1-1200
timeout 600 ncu --set full --clock-control none -k regex:k_matvec$ -c 24 -o gpurun_out/r02_07_k1 -f python tools/k1_sweep.py 16384 1 > gpurun_out/r02_07_ncu_k1.log 2>&1; echo "ncu k1 rc=$?"; tail -2 gpurun_out/r02_07_ncu_k1.log | cut -c1-300
