#!/bin/bash
# r2d: full GPU test suite (incl. PRT_FLAG_DIAGNOSE and the FP32 fast mode), bench line of config 4,
# kernel timings of both precisions, ncu captures of trace_kernel_f32 and of the bench's launch list
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r2d.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r2d.log
tail -15 $O/pytest_gpu_r2d.log
for cfg in "config4 16777216" "config5 33554432" "config2 100000" "config3 1048586"; do
  timeout 300 python scripts/kbench.py $cfg 2>&1 | grep -v "^$"
  KBENCH_PRECISION=fp32 timeout 300 python scripts/kbench.py $cfg 2>&1 | grep -v "^$" | sed 's/^default/fp32   /'
done | tee $O/kbench_r2d.txt
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_r2d_config4_1gpu.json 2> $O/bench_r2d_config4_1gpu.err; tail -c 1500 $O/bench_r2d_config4_1gpu.json
KBENCH_ONLY=k1 KBENCH_PRECISION=fp32 timeout 600 ncu --set full --import-source on --clock-control none \
   -k regex:trace_kernel_f32 -s 2 -c 1 -o $O/prof_trace_f32_r2d_config4 -f python scripts/kbench.py config4 16777216 > $O/ncu_f32_r2d.log 2>&1
tail -2 $O/ncu_f32_r2d.log
