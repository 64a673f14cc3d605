#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2c.log
tail -8 gpurun_out/pytest_gpu_r2c.log
V=pyrayt_b200/variants
for lib in $V/lib_r1_final.so $V/lib_v3.so $V/lib_v4.so $V/lib_v5.so; do
  for cfg in "config4 16777216" "config5 33554432"; do
    PYRAYT_B200_LIB=$lib timeout 300 python scripts/kbench.py $cfg 2>&1 | grep -v "^$"
  done
done | tee gpurun_out/kbench_r2c.txt
for cfg in config4; do
  n=16777216
  KBENCH_ONLY=k1 PYRAYT_B200_LIB=$V/lib_v5.so timeout 600 ncu --set full --import-source on --clock-control none \
     -k regex:trace_kernel -s 2 -c 1 -o gpurun_out/prof_trace_r2c_$cfg -f python scripts/kbench.py $cfg $n > gpurun_out/ncu_r2c_$cfg.log 2>&1
  tail -2 gpurun_out/ncu_r2c_$cfg.log
  KBENCH_ONLY=k1 PYRAYT_B200_LIB=$V/lib_v5.so timeout 600 ncu --set full --import-source on --clock-control none \
     -k regex:gather_kernel -s 2 -c 1 -o gpurun_out/prof_gather_r2c_$cfg -f python scripts/kbench.py $cfg $n > gpurun_out/ncu_gather_r2c_$cfg.log 2>&1
  tail -2 gpurun_out/ncu_gather_r2c_$cfg.log
done
