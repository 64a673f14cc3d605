#!/bin/bash
# Multi-GPU evidence on one 8-GPU box: NCCL parity test, strong scaling of config 4 (2^24 rays in
# total at 2/4/8 GPUs, e2e and the host copy ceiling included), config 5 at 8 x 2^25 rays.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/multi_r2s_gpus.txt; nproc >> $O/multi_r2s_gpus.txt; free -g | head -2 >> $O/multi_r2s_gpus.txt
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > $O/pytest_gpu_multi_r2s.log 2>&1; echo "rc=$?" >> $O/pytest_gpu_multi_r2s.log
tail -3 $O/pytest_gpu_multi_r2s.log
P=29510
# N = 1 on the same box (same host, same GPUs) so that the efficiencies below compare like with like
timeout 420 python bench.py --gpus 1 --steps 4 --warmup 3 --workload config4 --scaling strong --no-cpu \
  > $O/bench_r2s_config4_strong_1gpu.json 2> $O/bench_r2s_config4_strong_1gpu.err
tail -c 600 $O/bench_r2s_config4_strong_1gpu.json | head -c 300; echo
for n in 2 4 8; do
  P=$((P+1))
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P \
    bench.py --gpus $n --steps 4 --warmup 3 --workload config4 --scaling strong --no-cpu \
    > $O/bench_r2s_config4_strong_${n}gpu.json 2> $O/bench_r2s_config4_strong_${n}gpu.err
  tail -c 600 $O/bench_r2s_config4_strong_${n}gpu.json | head -c 300; echo
done
P=$((P+1))
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P \
  bench.py --gpus 8 --steps 4 --warmup 3 --workload config5 --scaling weak --no-cpu --no-e2e \
  > $O/bench_r2s_config5_weak_8gpu.json 2> $O/bench_r2s_config5_weak_8gpu.err
head -c 300 $O/bench_r2s_config5_weak_8gpu.json; echo
