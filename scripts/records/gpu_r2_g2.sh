#!/bin/bash
# Round 2, 2-GPU visit: BatchNorm reductions over NVLink peer memory: parity over NCCL ranks, then A/B inside the bench
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity.log gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_partition.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/test_gpu_partition.log 2>&1
echo "test_gpu_partition exit=$?" >> gpurun_out/summary.txt; tail -n 15 gpurun_out/test_gpu_partition.log
for mode in 1 0; do
  DDMP_PEER_ALLREDUCE=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$mode \
    bench.py --gpus 2 --steps 8 --warmup 3 --no-mode-a --no-e2e > gpurun_out/bench_2gpu_peer$mode.json 2> gpurun_out/bench_2gpu_peer$mode.err
  echo "bench 2gpu peer=$mode exit=$?" >> gpurun_out/summary.txt; cat gpurun_out/bench_2gpu_peer$mode.json; tail -n 4 gpurun_out/bench_2gpu_peer$mode.err
done
cat gpurun_out/summary.txt
