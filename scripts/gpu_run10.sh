#!/bin/bash
# r2l: equal-key detection compiled out of the closed-form merges (default) vs always on (lib_ties)
mkdir -p gpurun_out
O=gpurun_out
V=pyrayt_b200/variants
{
for rep in 1 2; do
for cfg in "config4 16777216" "config5 33554432"; do
  timeout 300 python scripts/kbench.py $cfg 2>&1 | grep -v "^$" | grep -v "record=none" | sed 's/^default/no ties  /'
  PYRAYT_B200_LIB=$V/lib_ties.so timeout 300 python scripts/kbench.py $cfg 2>&1 | grep -v "^$" | grep -v "record=none"
done
done
} | tee $O/kbench_r2l.txt
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu_r2l.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r2l.log; tail -3 $O/pytest_gpu_r2l.log
