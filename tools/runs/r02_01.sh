#!/bin/bash
# round 2, GPU visit 1: parity suite with the new large-golden / explicit-variant tests, the experimental paired
# back-transformation, and the bench line with the full-size quality block
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/r02_01_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_01_pytest.log | cut -c1-400
ZQ_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_zquatev.py -q -m gpu -k experimental > gpurun_out/r02_01_experimental.log 2>&1; echo "experimental rc=$?"; tail -3 gpurun_out/r02_01_experimental.log | cut -c1-400
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r02_01_bench.json 2> gpurun_out/r02_01_bench.err; echo "bench rc=$?"; grep '^{' gpurun_out/r02_01_bench.json | cut -c1-6000; tail -5 gpurun_out/r02_01_bench.err
ZQ_BT_PAIR=1 timeout 300 python tools/probe_solve.py 16384 2>&1 | tail -3 | cut -c1-600 | tee gpurun_out/r02_01_btpair.log
