#!/bin/bash
# r2q: K1 at three blocks per SM (80 registers, ~470 bytes of spills) against the default two (128 registers)
mkdir -p gpurun_out
{
for cfg in "config4 16777216" "config5 33554432"; do
  timeout 300 python scripts/kbench.py $cfg 2>&1 | grep -v "^$" | grep "record=all" | sed 's/^default/2 blocks, 128 regs/'
  PYRAYT_B200_LIB=pyrayt_b200/variants/lib_b3.so timeout 300 python scripts/kbench.py $cfg 2>&1 | grep -v "^$" | grep "record=all"
done
} | tee gpurun_out/kbench_r2q.txt
