#!/bin/bash
# r2i: FP32 ray-ordered traversal (3 vs 4 blocks per SM), GPU tests, compute-sanitizer on the extended exercise
mkdir -p gpurun_out
O=gpurun_out
V=pyrayt_b200/variants
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu_r2i.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r2i.log
tail -4 $O/pytest_gpu_r2i.log
{
for cfg in "config4 16777216" "config5 33554432" "config3 1048586"; do
  KBENCH_PRECISION=fp32 timeout 300 python scripts/kbench.py $cfg 2>&1 | grep -v "^$" | sed 's/^default/fp32 ordered 4 blocks/'
  KBENCH_PRECISION=fp32 PYRAYT_B200_LIB=$V/lib_f32b3.so timeout 300 python scripts/kbench.py $cfg 2>&1 | grep -v "^$" | sed 's/^pyrayt_b200.variants.lib_f32b3.so/fp32 ordered 3 blocks/'
done
} | tee $O/kbench_r2i.txt
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool python scripts/sanitize_run.py > $O/sanitizer_r2i_$tool.log 2>&1
  echo "[$tool] $(grep -E 'sanitize_run:|ERROR SUMMARY|RACECHECK SUMMARY|Error' $O/sanitizer_r2i_$tool.log | tr '\n' ' ')"
done | tee $O/sanitizer_r2i.txt
