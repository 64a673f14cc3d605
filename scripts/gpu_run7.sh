#!/bin/bash
# r2g: A/B -- FP32 kernel at 3 vs 4 blocks per SM; FP64 ray-ordered vs list-order traversal on the small scenes
mkdir -p gpurun_out
O=gpurun_out
V=pyrayt_b200/variants
{
for cfg in "config4 16777216" "config5 33554432"; do
  KBENCH_PRECISION=fp32 timeout 300 python scripts/kbench.py $cfg 2>&1 | grep -v "^$" | sed 's/^default/fp32 3 blocks/'
  KBENCH_PRECISION=fp32 PYRAYT_B200_LIB=$V/lib_f32b4.so timeout 300 python scripts/kbench.py $cfg 2>&1 | grep -v "^$" | sed 's/^pyrayt_b200.variants.lib_f32b4.so/fp32 4 blocks/'
done
for cfg in "config5 33554432" "config2 100000" "config3 1048586"; do
  timeout 300 python scripts/kbench.py $cfg 2>&1 | grep -v "^$" | sed 's/^default/fp64 list   /'
  PRT_FORCE_TRAVERSAL=ordered timeout 300 python scripts/kbench.py $cfg 2>&1 | grep -v "^$" | sed 's/^default/fp64 ordered/'
done
} | tee $O/kbench_r2g.txt
