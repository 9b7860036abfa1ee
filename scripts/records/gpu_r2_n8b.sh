#!/bin/bash
# Round 2, second 8-GPU visit: the driver's scaling command at N=8 with the BatchNorm reductions over NVLink peer memory
set -u
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
T0=$(date +%s)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_8gpu_weak_peer.json 2> gpurun_out/bench_8gpu_weak_peer.err
echo "bench 8gpu weak peer exit=$? wall=$(( $(date +%s) - T0 ))s" >> gpurun_out/summary.txt; cat gpurun_out/bench_8gpu_weak_peer.json; tail -n 5 gpurun_out/bench_8gpu_weak_peer.err
cat gpurun_out/summary.txt
