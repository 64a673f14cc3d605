#!/bin/bash
# kernel variant A/B + ncu captures (config 4 and 5) of the current library
mkdir -p gpurun_out
V=pyrayt_b200/variants
for lib in $V/lib_r1_final.so $V/lib_v2.so $V/lib_v2_t192.so; do
  for cfg in "config4 16777216" "config5 33554432"; do
    PYRAYT_B200_LIB=$lib timeout 300 python scripts/kbench.py $cfg 2>&1 | grep -v "^$"
  done
done | tee gpurun_out/kbench_r2b.txt
for cfg in config4 config5; do
  n=16777216; [ $cfg = config5 ] && n=33554432
  KBENCH_ONLY=k1 PYRAYT_B200_LIB=$V/lib_v2.so timeout 600 ncu --set full --import-source on --clock-control none \
     -k regex:trace_kernel -s 2 -c 1 -o gpurun_out/prof_trace_r2b_$cfg -f python scripts/kbench.py $cfg $n > gpurun_out/ncu_r2b_$cfg.log 2>&1
  tail -3 gpurun_out/ncu_r2b_$cfg.log
done
ls -la gpurun_out/*.ncu-rep | tail -3
