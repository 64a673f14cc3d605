#!/bin/bash
# Evidence run of a build: GPU tests, every config's bench line, the reference arm, the launch list of the
# bench command, ncu --set full of the trace kernels and the ordering pass, compute-sanitizer.
#   bash scripts/gpu_run_final.sh r2h
TAG=${1:-r2h}
mkdir -p gpurun_out
O=gpurun_out
{ nproc; free -g | head -2; nvidia-smi -L; ls baseline/_ref; } > $O/host_$TAG.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
tail -4 $O/pytest_gpu_$TAG.log
for wl in config4 config1 config2 config3 config5; do
  timeout 900 python bench.py --workload $wl > $O/bench_${TAG}_${wl}_1gpu.json 2> $O/bench_${TAG}_${wl}_1gpu.err
  echo "bench $wl rc=$?"; head -c 300 $O/bench_${TAG}_${wl}_1gpu.json; echo
done
timeout 900 python bench.py --impl reference > $O/bench_${TAG}_reference_arm.json 2> $O/bench_${TAG}_reference_arm.err
head -c 400 $O/bench_${TAG}_reference_arm.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $O/launches_$TAG.log 2>&1
n=16777216
KBENCH_ONLY=k1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:trace_kernel -s 2 -c 1 \
  -o $O/prof_trace_${TAG}_config4 -f python scripts/kbench.py config4 $n > $O/ncu_trace_$TAG.log 2>&1
KBENCH_ONLY=k1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:gather_kernel -s 2 -c 1 \
  -o $O/prof_gather_${TAG}_config4 -f python scripts/kbench.py config4 $n > $O/ncu_gather_$TAG.log 2>&1
KBENCH_ONLY=k1 KBENCH_PRECISION=fp32 timeout 600 ncu --set full --import-source on --clock-control none -k regex:trace_kernel_f32 -s 2 -c 1 \
  -o $O/prof_trace_f32_${TAG}_config4 -f python scripts/kbench.py config4 $n > $O/ncu_trace_f32_$TAG.log 2>&1
ls -la $O/*.ncu-rep | tail -4
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool python scripts/sanitize_run.py > $O/sanitizer_${TAG}_$tool.log 2>&1
  echo "[$tool] $(grep -E 'sanitize_run:|ERROR SUMMARY|RACECHECK SUMMARY' $O/sanitizer_${TAG}_$tool.log | tr '\n' ' ')"
done | tee $O/sanitizer_$TAG.txt
