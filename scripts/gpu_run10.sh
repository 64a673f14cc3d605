#!/bin/bash
# r2m: grouped exponent guards (integer min/max, default) vs per-value guards (lib_oldguards)
mkdir -p gpurun_out
O=gpurun_out
V=pyrayt_b200/variants
{
for rep in 1 2; do
for cfg in "config4 16777216" "config5 33554432"; do
  timeout 300 python scripts/kbench.py $cfg 2>&1 | grep -v "^$" | grep -v "record=none" | sed 's/^default/new guards/'
  PYRAYT_B200_LIB=$V/lib_oldguards.so timeout 300 python scripts/kbench.py $cfg 2>&1 | grep -v "^$" | grep -v "record=none"
done
done
} | tee $O/kbench_r2m.txt
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu_r2m.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r2m.log; tail -3 $O/pytest_gpu_r2m.log
