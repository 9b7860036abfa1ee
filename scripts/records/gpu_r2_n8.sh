#!/bin/bash
# Round 2, 8-GPU visit: the driver's scaling command at N=8 (one ~8M-face mesh partitioned over 8 GPUs, weak scaling),
# then the 16M-face config (BASELINE configs[4], strong scaling point at 8 GPUs)
set -u
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/gpus8.txt
(nproc; free -g) > gpurun_out/host8.txt 2>&1
T0=$(date +%s)
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_8gpu_weak.json 2> gpurun_out/bench_8gpu_weak.err
echo "bench 8gpu weak exit=$? wall=$(( $(date +%s) - T0 ))s" >> gpurun_out/summary.txt; cat gpurun_out/bench_8gpu_weak.json; tail -n 5 gpurun_out/bench_8gpu_weak.err
T0=$(date +%s)
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 \
  bench.py --gpus 8 --steps 5 --warmup 3 --mode partition --freq 895 --no-e2e > gpurun_out/bench_8gpu_16M.json 2> gpurun_out/bench_8gpu_16M.err
echo "bench 8gpu 16M exit=$? wall=$(( $(date +%s) - T0 ))s" >> gpurun_out/summary.txt; cat gpurun_out/bench_8gpu_16M.json; tail -n 5 gpurun_out/bench_8gpu_16M.err
cat gpurun_out/summary.txt
