#!/bin/bash
# Drop-in proof: the reference's unmodified test.cc, (a) linked with the reference objects, (b) linked with
# libzquatev_b200.so.  Both print max eigenvalue deviation vs zheev and ||V^H M V - L||^2 (test.cc:113-118).
mkdir -p gpurun_out
{
for n in 200 500 1000; do
  echo "=== test_ref.x $n (reference, $(nproc) host threads)"; OPENBLAS_NUM_THREADS=$(nproc) ./oracle/_ref/test_ref.x $n
  echo "=== test_b200.x $n (same test.cc linked against libzquatev_b200.so)"; OPENBLAS_NUM_THREADS=$(nproc) ./oracle/_ref/test_b200.x $n
done
} > gpurun_out/dropin.txt 2>&1
cat gpurun_out/dropin.txt
