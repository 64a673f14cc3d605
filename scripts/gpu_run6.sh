#!/bin/bash
# r2f: GPU tests, K1 timings after the zero-numerator shortcut, ncu source-level capture of K1 on config 4
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu_r2f.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r2f.log
tail -5 $O/pytest_gpu_r2f.log
for cfg in "config4 16777216" "config5 33554432"; do
  timeout 300 python scripts/kbench.py $cfg 2>&1 | grep -v "^$"
done | tee $O/kbench_r2f.txt
KBENCH_ONLY=k1 timeout 900 ncu --set full --import-source on --clock-control none \
   -k regex:trace_kernel -s 2 -c 1 -o $O/prof_trace_r2f_config4 -f python scripts/kbench.py config4 16777216 > $O/ncu_r2f_config4.log 2>&1
tail -2 $O/ncu_r2f_config4.log
