#!/bin/bash
# r2e: full GPU test suite again (FP32 thresholds, 973-leaf scene), ncu source-level capture of K1 on config 5
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu_r2e.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r2e.log
tail -15 $O/pytest_gpu_r2e.log
KBENCH_ONLY=k1 timeout 900 ncu --set full --import-source on --clock-control none \
   -k regex:trace_kernel -s 2 -c 1 -o $O/prof_trace_r2e_config5 -f python scripts/kbench.py config5 8388608 > $O/ncu_r2e_config5.log 2>&1
tail -2 $O/ncu_r2e_config5.log
