#!/bin/bash
# Round-2 first GPU call: full gpu test suite (incl. the live drop-in against baseline/_ref), every config's bench line,
# the reference arm.
set -x
mkdir -p gpurun_out
nproc > gpurun_out/host_r2.txt; free -g >> gpurun_out/host_r2.txt; nvidia-smi -L >> gpurun_out/host_r2.txt
ls baseline/_ref >> gpurun_out/host_r2.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2a.log
tail -5 gpurun_out/pytest_gpu_r2a.log
for wl in config4 config1 config2 config3 config5; do
  timeout 600 python bench.py --workload $wl > gpurun_out/bench_r2a_$wl.json 2> gpurun_out/bench_r2a_$wl.err
  echo "bench $wl rc=$?"; tail -c 600 gpurun_out/bench_r2a_$wl.json
done
timeout 600 python bench.py --impl reference > gpurun_out/bench_r2a_reference.json 2> gpurun_out/bench_r2a_reference.err
tail -c 1500 gpurun_out/bench_r2a_reference.json
