#!/bin/bash
# Round 2, final build on 2 GPUs: the NCCL / peer-memory parity test of the partitioned mode, then the driver's scaling
# command at N=2 (mode B weak scaling: one 2M-face mesh over 2 GPUs).
set -u
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_partition.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/test_partition_2gpu.log 2>&1
echo "test partition 2gpu exit=$? $(tail -n 1 gpurun_out/test_partition_2gpu.log)" >> gpurun_out/summary.txt
T0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_weak.json 2> gpurun_out/bench_2gpu_weak.err
echo "bench 2gpu weak exit=$? wall=$(( $(date +%s) - T0 ))s" >> gpurun_out/summary.txt
tail -n 8 gpurun_out/test_partition_2gpu.log; cat gpurun_out/bench_2gpu_weak.json; tail -n 5 gpurun_out/bench_2gpu_weak.err; cat gpurun_out/summary.txt
