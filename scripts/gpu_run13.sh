#!/bin/bash
# r2o: ordering pass variants (blocks per SM x quotient form), batched loads
mkdir -p gpurun_out
O=gpurun_out
V=pyrayt_b200/variants
{
for rep in 1 2; do
for lib in lib_gB4div lib_gB3div lib_gB4tilt lib_gB5div; do
  PYRAYT_B200_LIB=$V/$lib.so timeout 300 python scripts/kbench.py config4 16777216 2>&1 | grep -v "^$" | grep "whole step"
done
done
PYRAYT_B200_LIB=$V/lib_gB4div.so timeout 300 python scripts/kbench.py config5 33554432 2>&1 | grep -v "^$" | grep "whole step\|record=all"
PYRAYT_B200_LIB=$V/lib_gB4tilt.so timeout 300 python scripts/kbench.py config5 33554432 2>&1 | grep -v "^$" | grep "whole step\|record=all"
} | tee $O/kbench_r2o.txt
