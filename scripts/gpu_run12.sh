#!/bin/bash
# r2n: ordering pass back at four blocks per SM (61 registers, streamed columns) vs five blocks; GPU tests
mkdir -p gpurun_out
O=gpurun_out
V=pyrayt_b200/variants
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu_r2n.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r2n.log; tail -3 $O/pytest_gpu_r2n.log
{
for rep in 1 2; do
for cfg in "config4 16777216" "config5 33554432"; do
  timeout 300 python scripts/kbench.py $cfg 2>&1 | grep -v "^$" | grep -v "record=none" | sed 's/^default/gather 4 blocks/'
  PYRAYT_B200_LIB=$V/lib_gather5.so timeout 300 python scripts/kbench.py $cfg 2>&1 | grep -v "^$" | grep -v "record=none"
done
done
} | tee $O/kbench_r2n.txt
