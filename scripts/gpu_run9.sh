#!/bin/bash
# r2j: Cube-leaf slab test A/B (A = guarded generic form, B = guard-free form inlined, C = guard-free form out of line)
mkdir -p gpurun_out
O=gpurun_out
V=pyrayt_b200/variants
{
for lib in lib_cubeA lib_cubeB lib_cubeC; do
for cfg in "config4 16777216" "config5 33554432"; do
  PYRAYT_B200_LIB=$V/$lib.so timeout 300 python scripts/kbench.py $cfg 2>&1 | grep -v "^$" | grep -v "record=none"
done
done
KBENCH_PRECISION=fp32 timeout 300 python scripts/kbench.py config5 33554432 2>&1 | grep -v "^$" | sed 's/^default/fp32 list-order again/'
KBENCH_PRECISION=fp32 timeout 300 python scripts/kbench.py config4 16777216 2>&1 | grep -v "^$" | sed 's/^default/fp32 ordered/'
} | tee $O/kbench_r2j.txt
