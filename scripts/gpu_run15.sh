#!/bin/bash
# r2t: denominators' exponent windows folded into the grouped guards (default) vs the previous build (lib_prev)
mkdir -p gpurun_out
{
for rep in 1 2; do
for cfg in "config4 16777216" "config5 33554432"; do
  timeout 300 python scripts/kbench.py $cfg 2>&1 | grep -v "^$" | grep -v "record=none" | sed 's/^default/fused guards/'
  PYRAYT_B200_LIB=pyrayt_b200/variants/lib_prev.so timeout 300 python scripts/kbench.py $cfg 2>&1 | grep -v "^$" | grep -v "record=none"
done
done
} | tee gpurun_out/kbench_r2t.txt
