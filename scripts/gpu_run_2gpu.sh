#!/bin/bash
# two-GPU sanity of the final build: the whole GPU suite (the NCCL test is not skipped here) and the bench at N = 2
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu_r2k_2gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r2k_2gpu.log
tail -4 $O/pytest_gpu_r2k_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_r2k_config4_weak_2gpu.json 2> $O/bench_r2k_config4_weak_2gpu.err
tail -c 400 $O/bench_r2k_config4_weak_2gpu.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 \
  bench.py --gpus 2 --steps 3 --warmup 3 --impl reference > $O/bench_r2k_reference_arm_2gpu.json 2> $O/bench_r2k_reference_arm_2gpu.err
head -c 300 $O/bench_r2k_reference_arm_2gpu.json; echo
